"""Debug aid: per-rank, per-step mismatch report of the slab path against the single-GPU run (torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import interfaceadvection.jl_b200 as ia
from interfaceadvection.jl_b200 import configs, slab

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for dtype, per_z, N, hook in [("float32", False, (96, 64, 16), True), ("float64", False, (96, 64, 16), True), ("float32", False, (96, 64, 16), False),
                                  ("float32", True, (96, 64, 16), True)]:
        perdir = (1, 2, 3) if per_z else (1, 2)
        T = getattr(torch, dtype); N1, N2, nz = N; Ng = (N1, N2, nz * world)
        case = configs.make_case(Ng, dtype=dtype, device=dev, kind="C4", vel="enright")
        # sphere straddling the boundary between ranks 0 and 1 AND reaching rank 2
        sdf = configs.sdf_sphere([N1 / 2, N2 / 2, nz * 1.5], nz * 0.9, inside_dark=False)
        sim = ia.TwoPhaseSimulation(Ng, (0, 0, 0), float(N1), T=T, lam_rho=1e-3, InterfaceSDF=sdf, perdir=perdir, U=1.0, dt=1.0, device=dev)
        sim.flow.u.copy_(case["u"]); ia.BC(sim.flow.u, (0, 0, 0), False, perdir)
        g = slab.SlabGeom(rank, world, nz, slab.G_DEFAULT, per_z); nzg = nz * world
        zidx = torch.tensor([((g.z_origin + l - 1) % nzg) + 1 if per_z else min(max(g.z_origin + l, 0), nzg + 1) for l in range(g.nz_local + 2)], device=dev)
        run = slab.SlabRunner(N, dtype, perdir, "C4", rank, world, dev, fields=(sim.intf.f.index_select(2, zidx), sim.flow.u.index_select(2, zidx)))

        def project(a, c, stage):
            v = torch.empty_like(a.u)
            ia.rhou2u(v, c.rhou, c.f0 if stage == "predictor" else c.f, c.lam_rho)
            a.u[1:-1, 1:-1, 1:-1] = a.u[1:-1, 1:-1, 1:-1] * 0.9 + v[1:-1, 1:-1, 1:-1] * 0.1
            ia.BC(a.u, a.uBC, False, a.perdir)
        sl = slice(1 + rank * nz, 1 + (rank + 1) * nz)
        for n in range(3):
            run.step(project=project if hook else None)
            ia.mom_advect_step(sim.flow, sim.intf, 1.0, project=project if hook else None); sim.flow.dt.append(1.0)
            torch.cuda.synchronize()
            msg = []
            for name, ref, got in [("f", sim.intf.f[1:-1, 1:-1, sl], run.owned_f()), ("rhou", sim.intf.rhou[1:-1, 1:-1, sl, :], run.owned_rhou()),
                                   ("u", sim.flow.u[1:-1, 1:-1, sl, :], run.flow.u[1:-1, 1:-1, run.geom.owned, :]),
                                   ("f0", sim.intf.f0[1:-1, 1:-1, sl], run.intf.f0[1:-1, 1:-1, run.geom.owned])]:
                ne = (ref != got)
                if ne.any():
                    planes = ne.reshape(ne.shape[0] * ne.shape[1], ne.shape[2], -1).any(0).any(-1).nonzero().flatten().tolist()
                    msg.append(f"{name}: {int(ne.sum())} cells differ, local owned planes {planes}, max|d|={float((ref - got).abs().max()):.3e}")
            print(f"[dbg] {dtype} per_z={per_z} hook={hook} step {n} rank {rank}: " + ("; ".join(msg) if msg else "OK"), flush=True)
        dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
