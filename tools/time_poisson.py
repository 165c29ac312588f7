"""Times the pressure projection (update!, psolver!, myproject!) on cuda:0 with CUDA events: a two-phase problem
(sphere of radius n/4, density ratio 1/λρ = 1000), source = divergence of a random velocity, a FIXED number of Jacobi-PCG iterations
(the iteration count to convergence is a property of the problem, not of the kernels).  Per iteration the kernels move 17 s B per
cell (ifadv_poisson.cuh header): the line reports that algorithmic rate against the measured HBM peak.
`python tools/time_poisson.py [n dtype itmx]`; bench.py imports measure() for its `projection` sub-line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _peak():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6650 GB/s (of fallback)"


def measure(n=512, f64=False, itmx=50, dev="cuda", perdir=(), reps=3, cpu_n=0):
    import numpy as np
    import torch
    import interfaceadvection.jl_b200 as ia

    N = (n,) * 3
    T = torch.float64 if f64 else torch.float32
    s = 8 if f64 else 4
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), float(n), T=T, lam_rho=1e-3, dt=0.25, psolver="Poisson", perdir=perdir, device=dev,
                                InterfaceSDF=lambda x: ((x - n / 2) ** 2).sum(-1).sqrt() - n / 4)
    a, c, b = sim.flow, sim.intf, sim.pois
    gen = torch.Generator(device=dev).manual_seed(5)
    a.u.copy_(0.1 * torch.randn(a.u.shape, generator=gen, device=dev, dtype=T))
    ia.BC(a.u, a.uBC, False, perdir)
    ia.updateL(a.mu0, c.f, c.lam_rho, perdir, fill_one=True)
    ia.update(b)
    ctx = ia.context_for(c.f)
    src = torch.zeros_like(b.z)
    ins = (slice(1, -1),) * 3
    for i in range(3):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(3)) + (i,)
        src[ins] += a.u[hi] - a.u[ins + (i,)]
    res = []
    for _ in range(reps):
        b.z.copy_(src); b.x.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record()
        it = ia.psolver(b, itmx=itmx)
        e1.record(); torch.cuda.synchronize()
        res.append((it, e0.elapsed_time(e1), ctx.launches - l0))
    it, ms, launches = min(res, key=lambda r: r[1])
    # the whole projection once, to convergence (or the reference's cap of 2000 iterations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.p.zero_()
    e0.record()
    n_full = ia.myproject(a, b, 1.0)
    e1.record(); torch.cuda.synchronize()
    ms_full = e0.elapsed_time(e1)
    cells = float(np.prod(N))
    peak, src_peak = _peak()
    gbps = 17 * s * cells * it / (ms * 1e-3) / 1e9
    out = {"what": "pressure projection (SURVEY 8f row 2): Jacobi-PCG psolver! of src/flow.jl:300-326, sphere R = n/4 with density ratio 1000, "
                   "source = divergence of a random velocity; fixed iteration count",
           "grid": list(N), "dtype": "f64" if f64 else "f32", "perdir": list(perdir), "iterations": it, "ms": ms, "ms_per_iteration": ms / max(it, 1),
           "launches": launches, "bytes_per_cell_per_iteration": 17 * s, "algorithmic_GBps": gbps, "peak_GBps": peak, "peak_source": src_peak,
           "frac_of_hbm_roofline": gbps / peak, "Gcell_iterations_per_s": cells * it / (ms * 1e-3) / 1e9,
           "myproject_to_convergence": {"iterations": n_full, "ms": ms_full, "r2": b.r2[-1], "tol": 50 * float(torch.finfo(T).eps)}}
    if cpu_n:
        out["cpu_port"] = measure_cpu(cpu_n, f64, min(itmx, 20))
    return out


def measure_cpu(n, f64, itmx):
    """The oracle's psolver! (OpenMP build, all host threads) on the same kind of problem at n³ -- the CPU figure beside the GPU one."""
    import time
    import numpy as np
    from oracle import pyoracle as O

    T = np.float64 if f64 else np.float32
    Ng = (n + 2,) * 3
    O.set_num_threads(os.cpu_count() or 1)
    idx = np.indices(Ng).astype(np.float32)
    f = np.asfortranarray((((idx - (n / 2 + 1)) ** 2).sum(0) ** 0.5 < n / 4).astype(T))
    mu0 = np.asfortranarray(np.ones(Ng + (3,), T))
    O.updateL(mu0, f, 1e-3, (), omp=True)
    rng = np.random.default_rng(5)
    z = O.zeros(Ng, T)
    bsrc = rng.standard_normal((n,) * 3)
    z[1:-1, 1:-1, 1:-1] = (bsrc - bsrc.mean()).astype(T)
    p = O.Poisson(O.zeros(Ng, T), mu0, z, ())
    t0 = time.perf_counter()
    it, _ = O.psolver(p, itmx=itmx, omp=True)
    dt = time.perf_counter() - t0
    return {"grid": [n] * 3, "iterations": it, "ms_per_iteration": dt * 1e3 / max(it, 1), "cores": O.num_threads(),
            "Gcell_iterations_per_s": n ** 3 * it / dt / 1e9, "kind": "port (C++/OpenMP restatement, un-fused pass structure)"}


def measure_slab(N=(512, 512, 256), f64=False, itmx=50, reps=3):
    """Weak scaling of the solver over z-slabs (one process per GPU, torchrun): N per GPU, fixed iteration count, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import interfaceadvection.jl_b200 as ia
    from interfaceadvection.jl_b200 import slab

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = "float64" if f64 else "float32"
    T = getattr(torch, dtype)
    run = slab.SlabRunner(N, dtype, (), "C4", rank, world, dev, lam_rho=1e-3)
    a, c = run.flow, run.intf
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    a.u.copy_(0.1 * torch.randn(a.u.shape, generator=gen, device=dev, dtype=T))
    ia.BC(a.u, a.uBC, False, run.perdir)
    run.exchange(a.u)
    ia.updateL(a.mu0, c.f, c.lam_rho, run.perdir, fill_one=True)
    b = ia.Poisson(a.p, a.mu0, a.sigma, run.perdir)
    src = torch.zeros_like(b.z)
    ins = (slice(1, -1),) * 3
    for i in range(3):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(3)) + (i,)
        src[ins] += a.u[hi] - a.u[ins + (i,)]
    best = None
    for _ in range(reps):
        b.z.copy_(src); b.x.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        it = ia.psolver(b, itmx=itmx)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms) if best is None else min(best, float(ms))
    cells = float(np.prod(N)) * world
    s = 8 if f64 else 4
    if rank == 0:
        print(json.dumps({"what": "psolver on z-slabs, weak scaling", "n_gpus": world, "grid_per_gpu": list(N), "dtype": "f64" if f64 else "f32",
                          "iterations": it, "ms_per_iteration": best / max(it, 1), "Gcell_iterations_per_s": cells * it / (best * 1e-3) / 1e9,
                          "algorithmic_GBps_per_gpu": 17 * s * cells / world * it / (best * 1e-3) / 1e9,
                          "transport": run.transport, "r2": b.r2[-1]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "slab":
        measure_slab()
        sys.exit(0)
    if len(sys.argv) > 1:  # n dtype itmx: one configuration (for ncu)
        print(json.dumps(measure(int(sys.argv[1]), sys.argv[2] == "float64", int(sys.argv[3]), reps=1)), flush=True)
        sys.exit(0)
    for n, f64, itmx, per in ((64, False, 200, ()), (128, False, 200, ()), (256, False, 200, ()), (256, True, 200, ()), (512, False, 100, ()),
                              (512, False, 100, (1, 2))):
        print(json.dumps(measure(n, f64, itmx, perdir=per, cpu_n=128 if n == 512 and not per else 0)), flush=True)
