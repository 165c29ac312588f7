#!/usr/bin/env python
"""bench.py -- VOF+CMOM advection throughput of the B200-native path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one CMOM advection step = the transport half of MPFMomStep! (src/flow.jl:61,69-70,74,89-92):
2 x (u2ρu! + BC! + advectfq!) + the midpoint f⁰ = 2·D fused directional sweeps.  MPCFL and the
Poisson solve are excluded (timed separately by the reference's users).

Workload at N=1: BASELINE.json config 4, the 512³ Float32 rising bubble (periodic x,y; λρ=1e-3; Koren; WH
normals) with a synthetic solenoidal Taylor-Green velocity -- the grid north_star quotes its roofline target on.
N>1 (torchrun, one rank per GPU): the same 512x512x512 block per GPU, stacked along z (weak scaling, z-slabs with
ghost-plane exchange).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "VOF+CMOM advection Gcell-updates/s"
UNIT = "Gcell-updates/s"

WORKLOADS = {
    # name: (N per GPU, dtype, perdir, kind)
    "C4_bubble_512_f32": ((512, 512, 512), "float32", (1, 2), "C4"),
    "C3_dambreak_512x256x256_f32": ((512, 256, 256), "float32", (), "C3"),
    "C4_bubble_256_f32": ((256, 256, 256), "float32", (1, 2), "C4"),
    "C4_bubble_256_f64": ((256, 256, 256), "float64", (1, 2), "C4"),
    "C4_bubble_128_f32": ((128, 128, 128), "float32", (1, 2), "C4"),
    "C4_bubble_64_f32": ((64, 64, 64), "float32", (1, 2), "C4"),
    # pure-VOF configs (one step = advect! = D directional sweeps): reported for completeness, single GPU, no e2e / CPU legs
    "C2_enright_256_f32": ((256, 256, 256), "float32", (), "C2"),
    "C2_enright_256_f64": ((256, 256, 256), "float64", (), "C2"),
    "C1_zalesak_128_f64": ((128, 128), "float64", (), "C1"),
}
VOF_KINDS = ("C1", "C2")
CPU_SAMPLE = "C4_bubble_128_f32"  # bounded sample of the same workload for the CPU legs


def algorithmic_bytes_per_cell_sweep(D, s):
    return (2 * D + 7) * s + 1  # reads f,u_d,u⁰_d,ρu(D),uOld(D),c̄ ; writes f,ρu(D)


def algorithmic_bytes_per_cell_step(D, s):
    # 2 x [D sweeps + c̄ write + u2ρu! (2D+1)s]
    return 2 * (D * algorithmic_bytes_per_cell_sweep(D, s) + 1 + (2 * D + 1) * s)


def vof_bytes_per_cell_sweep(D, s):
    return 4 * s + 1  # reads f,u_d,u⁰_d,c̄ ; writes f   (the optional ρuf[·,d] output of advect! adds s)


def vof_bytes_per_cell_step(D, s):
    return D * vof_bytes_per_cell_sweep(D, s) + 1


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            # nvidia-smi needs 0.2-1 s before its first row: the timed region only starts once the stream is flowing
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        inside = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15]
        # a timed region shorter than the 100 ms sampling period may hold no row: take the rows next to it (the GPU has been
        # under the same load since the warm-up steps)
        rows = inside if inside else [r for r in self.rows if t0 - 0.5 <= r[0] <= t1 + 0.5]
        for t, line in rows:
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[1])); mx = float(p[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm, OpenMP build) -- the only place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------------------------
def cpu_case(name):
    from interfaceadvection.jl_b200 import configs
    from oracle import pyoracle as O

    N, dtype, perdir, kind = WORKLOADS[name]
    T = np.dtype(dtype).type
    case = configs.make_case(N, dtype=dtype, kind=kind)
    D = len(N)
    Ng = tuple(n + 2 for n in N)
    f = O.zeros(Ng, T); al = O.zeros(Ng, T); nh = O.zeros(Ng + (D,), T)
    O.applyVOF(f, al, nh, case["sdf"]); O.BCf(f, perdir)
    u = np.asfortranarray(case["u"].astype(T)); O.BC(u, (0,) * D, False, perdir)
    return dict(N=N, D=D, Ng=Ng, T=T, perdir=perdir, f=f, u=u, lam_rho=case["lam_rho"])


def cpu_step_fn(c):
    """One CMOM advection step with the reference's un-fused pass structure, all host threads."""
    from oracle import pyoracle as O

    T, D, Ng = c["T"], c["D"], c["Ng"]
    z = lambda *s: O.zeros(s, T)
    a = dict(ff=z(*Ng), alpha=z(*Ng), nhat=z(*Ng, D), cbar=np.zeros(Ng, dtype=np.int8, order="F"), rhou=z(*Ng, D), r=z(*Ng, D),
             Phi=z(*Ng), rhouf=z(*Ng, D), drho=z(*Ng, D), f0=z(*Ng), u0=z(*Ng, D))
    a["drho"][...] = 1
    state = {"n": 0}
    f, u, per, lr = c["f"], c["u"], c["perdir"], c["lam_rho"]

    def step():
        n = 1 + state["n"]
        dirO = tuple((n + i) % D + 1 for i in range(1, D + 1))
        a["u0"][...] = u; a["f0"][...] = f
        O.u2rhou(a["rhou"], a["u0"], a["f0"], lr, omp=True); O.BC(a["rhou"], (0,) * D, False, per)
        O.advectVOFrhouu(a["f0"], a["ff"], a["alpha"], a["nhat"], a["u0"], u, 1.0, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"],
                         a["nhat"], u, a["alpha"], a["drho"], lr, "Koren", "WH", (0,) * D, per, False, dirO, omp=True)
        a["f0"][...] = (a["f0"] + f) * T(0.5)
        a["f0"][...] = f
        O.u2rhou(a["rhou"], a["u0"], f, lr, omp=True); O.BC(a["rhou"], (0,) * D, False, per)
        O.advectVOFrhouu(f, a["ff"], a["alpha"], a["nhat"], u, u, 1.0, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"],
                         a["u0"], a["alpha"], a["drho"], lr, "Koren", "WH", (0,) * D, per, False, dirO, omp=True)
        state["n"] += 1
    return step


def time_cpu(name, steps, warmup):
    from oracle import pyoracle as O

    O.build()
    c = cpu_case(name)
    step = cpu_step_fn(c)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    cells = math.prod(c["N"])
    return cells * steps / dt / 1e9, dt / steps * 1e3, O.num_threads(True), c


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload or "C4_bubble_512_f32"
    N, dtype, perdir, kind = WORKLOADS[wl]
    val, ms, cores, c = time_cpu(CPU_SAMPLE, args.steps, args.warmup)
    sample = (f"{CPU_SAMPLE}: {'x'.join(map(str, c['N']))} {dtype} sub-grid of the workload (same SDF/velocity generators scaled to the box), "
              f"{args.steps} steps; restated reference algorithm (C++/OpenMP, un-fused pass structure), not the Julia package; "
              f"{cores} threads on {cpu_model()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64",
        "data": "synthetic",
        "config": {"workload": wl, "grid_per_gpu": list(N), "perdir": list(perdir), "limiter": "Koren", "normal_scheme": "WH",
                   "lambda_rho": 1e-3, "measured_on": "bounded CPU sample " + "x".join(map(str, c["N"]))},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import interfaceadvection.jl_b200 as ia
    from interfaceadvection.jl_b200 import configs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload or "C4_bubble_512_f32"
    N, dtype, perdir, kind = WORKLOADS[wl]
    D = len(N)
    T = getattr(torch, dtype)
    s = 4 if dtype == "float32" else 8
    vof = kind in VOF_KINDS
    if vof:
        if world > 1:
            raise SystemExit("the pure-VOF workloads are single-GPU report lines")
        args.no_e2e = args.no_cpu = True

    if world > 1:
        from interfaceadvection.jl_b200 import slab
        runner = slab.SlabRunner(N, dtype, perdir, kind, rank, world, dev)
    else:
        case = configs.make_case(N, dtype=dtype, device=dev, kind=kind)
        sim = ia.TwoPhaseSimulation(N, (0,) * D, float(N[0]), T=T, lam_rho=case["lam_rho"], InterfaceSDF=case["sdf"], perdir=perdir, U=1.0,
                                    dt=1.0, device=dev)
        sim.flow.u.copy_(case["u"])
        ia.BC(sim.flow.u, (0,) * D, False, perdir)
        sim.flow.u0.copy_(sim.flow.u)
        del case
        ctx = ia.context_for(sim.intf.f)

        class _Single:
            def __init__(self):
                self.n = 0

            def step(self):
                if vof:
                    ia.advect(sim.flow, sim.intf, check=False)  # advect!(a,c): pure VOF with u⁰, u (advection.jl:17-23)
                else:
                    ia.mom_advect_step(sim.flow, sim.intf, 1.0)
                sim.flow.dt.append(1.0)  # fixed Δt; advances the sweep-order rotation like push!(Δt) would

            contexts = [ctx]

            def mass(self):
                return ia.sum_inside(sim.intf.f)
        runner = _Single()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    m0 = runner.mass()
    for _ in range(args.warmup):
        runner.step()
    barrier()
    l0 = sum(c.launches for c in runner.contexts)
    for c in runner.contexts:
        c.profile(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        runner.step()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    kms, kn, fms, fn_ = 0.0, 0, 0.0, 0  # standard sweeps / fused first sweeps
    for c in runner.contexts:
        (a, b), (a1, b1) = c.profile_read()
        kms += a; kn += b; fms += a1; fn_ += b1
        per_dir = c.profile_read_dirs()
        c.profile(False)
    launches = sum(c.launches for c in runner.contexts) - l0
    if dist is not None:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    m1 = runner.mass()
    cells_gpu = math.prod(N)
    cells = cells_gpu * world
    value = cells * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the fused directional sweep) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    bytes_launch = (vof_bytes_per_cell_sweep(D, s) if vof else algorithmic_bytes_per_cell_sweep(D, s)) * cells_gpu
    avg_ms = kms / max(kn, 1)
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9 if kn else None
    # fused first sweeps (u2rhou + BC folded in): 3 fewer input streams -> (2D+4)s+1 B per cell
    fbytes = ((2 * D + 4) * s + 1) * cells_gpu
    favg = fms / max(fn_, 1)
    fused_info = {"avg_launch_ms": favg, "launches_timed": fn_, "algorithmic_bytes_per_launch": fbytes,
                  "achieved_gbs": (fbytes / (favg * 1e-3) / 1e9) if fn_ else None} if fn_ else None
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(wl, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "kernel": "ifadv::along2_kernel (y, z) / ifadv::xsweep_kernel (x): fused VOF+CMOM directional sweep, standard 13s+1 B/cell form, average over the three directions", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_launch, "avg_launch_ms": avg_ms, "launches_timed": kn,
                "sweep_share_of_step": ((kms + fms) / ms) if ms else None, "fused_first_sweep": fused_info,
                "ms_per_launch_by_direction": per_dir,
                "step_frac_of_roofline": ((vof_bytes_per_cell_step(D, s) if vof else algorithmic_bytes_per_cell_step(D, s)) * cells_gpu
                                          / (ms / args.steps * 1e-3) / 1e9) / peak}
    if vof:
        roofline["kernel"] = "fused pure-VOF directional sweep (ifadv::along2_kernel y/z, ifadv::march_kernel x; 2-D: ifadv::sweep_kernel), 4s+1 B/cell"
        roofline["fused_first_sweep"] = None

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D/D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        if world == 1:
            e2e = run_e2e(ia, torch, N, dtype, perdir, kind, dev, min(args.steps, args.e2e_steps))
        else:
            del runner
            torch.cuda.empty_cache()
            e2e = run_e2e_slabs(ia, torch, dist, N, dtype, perdir, kind, dev, rank, world, min(args.steps, args.e2e_steps))

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cms, cores, c = time_cpu(CPU_SAMPLE, args.cpu_steps, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE}: {'x'.join(map(str, c['N']))} {dtype} sub-grid of the workload, {args.cpu_steps} steps, {cms:.0f} ms/step; "
                         f"restated reference algorithm (C++/OpenMP, un-fused pass structure), not the Julia package; {cpu_model()}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": wl, "grid_per_gpu": list(N), "global_grid": [N[0], N[1], N[2] * world] if D == 3 else list(N),
                       "perdir": list(perdir), "limiter": "Koren", "normal_scheme": "WH", "lambda_rho": 1e-3,
                       "step": "pure-VOF advection step = advect! = D fused sweeps" if vof else "CMOM advection step = 2 x (u2rhou + BC + advectfq) + midpoint f0 = 6 fused sweeps (u2rhou+BC folded into the first sweep of each group); the u0<-u / f0<-f copies and the midpoint run on a second stream underneath the sweeps",
                       "l2": "working set >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                       "mass_drift_rel": abs(m1 - m0) / abs(m0) if m0 else None},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_e2e(ia, torch, N, dtype, perdir, kind, dev, steps):
    """Same metric through ifadv_mom_advect_step_host: pinned host f,u -> device, one CMOM step, f,ρu -> host."""
    from interfaceadvection.jl_b200 import configs

    D = len(N)
    T = getattr(torch, dtype)
    Ng = tuple(n + 2 for n in N)
    case = configs.make_case(N, dtype=dtype, device=dev, kind=kind)
    f = ia.jl_zeros(Ng, T, dev); al = ia.jl_zeros(Ng, T, dev); nh = ia.jl_zeros(Ng + (D,), T, dev)
    ia.applyVOF(f, al, nh, case["sdf"]); ia.BCf(f, perdir)
    u = case["u"]; ia.BC(u, (0,) * D, False, perdir)
    # pinned host buffers holding the column-major bytes
    fh = torch.empty(tuple(reversed(Ng)), dtype=T, pin_memory=True)
    uh = torch.empty((D,) + tuple(reversed(Ng)), dtype=T, pin_memory=True)
    rh = torch.empty((D,) + tuple(reversed(Ng)), dtype=T, pin_memory=True)
    fh.copy_(f.permute(*reversed(range(D)))); uh.copy_(u.permute(*reversed(range(D + 1))))
    del f, al, nh, u, case
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    ctx = ia.Context(Ng, dtype, dev.index or 0)
    lim, ns = ia.LIMITERS["Koren"], ia.NORMAL_SCHEMES["WH"]

    def dirO(n):
        return tuple((1 + n + i) % D + 1 for i in range(1, D + 1))
    ctx.mom_advect_step_host(fh.data_ptr(), uh.data_ptr(), rh.data_ptr(), 1.0, 1e-3, lim, ns, (0,) * D, perdir, dirO(0))  # warm-up + alloc
    ctx.mom_advect_step_host(fh.data_ptr(), uh.data_ptr(), rh.data_ptr(), 1.0, 1e-3, lim, ns, (0,) * D, perdir, dirO(1))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(steps):
        ctx.mom_advect_step_host(fh.data_ptr(), uh.data_ptr(), rh.data_ptr(), 1.0, 1e-3, lim, ns, (0,) * D, perdir, dirO(2 + n))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    esz = fh.element_size()
    S = math.prod(Ng)
    h2d, d2h, slabs = ctx.host_step_bytes()  # what the last call really copied (the overlap planes of the z-slab pipeline included)
    assert d2h >= (1 + D) * S * esz
    out = {"value": math.prod(N) * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": dt / steps * 1e3, "steps": steps,
           "api": "ifadv_mom_advect_step_host (C ABI, pinned host buffers: f,u in; f,rhou out)",
           "pipeline": f"{slabs} z-slabs, H2D / step / D2H of consecutive slabs overlap on three streams (8 overlap planes per interior slab end, "
                       "bit-identical to the single pass)" if slabs > 1 else "single pass",
           "checksum_f": float(fh.double().sum().item())}
    ctx.close()
    return out


def run_e2e_slabs(ia, torch, dist, N, dtype, perdir, kind, dev, rank, world, steps):
    """The e2e leg at N > 1 GPUs: every rank keeps ITS z-slab of the global state (owned planes + 8 overlap planes per interior
    end) in pinned host memory and advances it with ifadv_mom_advect_step_host; after each step the overlap planes of the host
    f are refreshed from the neighbours' owned planes (staged through the device, NCCL send/recv) -- the host-level halo
    exchange a multi-process caller has to do.  Timed region: barrier, steps x (host step + exchange), barrier; max over ranks."""
    from interfaceadvection.jl_b200 import slab

    T = getattr(torch, dtype)
    r = slab.SlabRunner(N, dtype, perdir, kind, rank, world, dev)  # fresh initial state of this rank's slab
    g = r.geom
    Ngl = tuple(r.intf.f.shape)
    fdev = r.intf.f
    # every rank pins 7 fields of its slab (~4 GB at 512^3): if one rank cannot, all ranks skip the leg together
    ok = torch.ones(1, device=dev)
    try:
        fh = torch.empty(tuple(reversed(Ngl)), dtype=T, pin_memory=True)
        uh = torch.empty((3,) + tuple(reversed(Ngl)), dtype=T, pin_memory=True)
        rh = torch.empty((3,) + tuple(reversed(Ngl)), dtype=T, pin_memory=True)
    except Exception as e:  # noqa: BLE001
        print(f"[bench] rank {rank}: cannot pin the host buffers of the e2e leg: {e}", file=sys.stderr)
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        return None
    fh.copy_(fdev.permute(2, 1, 0)); uh.copy_(r.flow.u.permute(3, 2, 1, 0))
    lperdir = r.perdir
    del r.flow, r.intf.rhou, r.intf.rhouf
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    ctx = ia.Context(Ngl, dtype, dev.index or 0)
    lim, ns = ia.LIMITERS["Koren"], ia.NORMAL_SCHEMES["WH"]
    o, W = g.owned, g.W
    fdz = fdev.permute(2, 1, 0)  # (z,y,x) view: a z-range is one contiguous block, like fh
    xbytes = [0, 0]

    def host_exchange():
        # owned edge planes host -> device, neighbours swap them (one NCCL batch), overlap planes device -> host
        for z0, z1 in ((o.start, o.start + W), (o.stop - W, o.stop)):
            fdz[z0:z1].copy_(fh[z0:z1], non_blocking=True)
            xbytes[0] += fh[z0:z1].numel() * fh.element_size()
        slab.exchange_overlap([fdev], g)
        for z0, z1 in ((o.start - g.wlo, o.start), (o.stop, o.stop + g.whi)):
            if z1 > z0:
                fh[z0:z1].copy_(fdz[z0:z1], non_blocking=True)
                xbytes[1] += fh[z0:z1].numel() * fh.element_size()
        torch.cuda.synchronize()

    def dirO(n):
        return tuple((1 + n + i) % 3 + 1 for i in range(1, 4))

    def step(n):
        ctx.mom_advect_step_host(fh.data_ptr(), uh.data_ptr(), rh.data_ptr(), 1.0, 1e-3, lim, ns, (0, 0, 0), lperdir, dirO(n))
        host_exchange()
    step(0); step(1)
    xbytes[0] = xbytes[1] = 0
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(steps):
        step(2 + n)
    dist.barrier(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    h2d, d2h, slabs = ctx.host_step_bytes()
    tot = torch.tensor([h2d + xbytes[0] // steps, d2h + xbytes[1] // steps], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    chk = fh[o].double().sum().to(dev)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    out = {"value": math.prod(N) * world * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
           "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": dt / steps * 1e3, "steps": steps,
           "api": "ifadv_mom_advect_step_host per rank on its z-slab (C ABI, pinned host buffers) + host-level overlap exchange of f "
                  "(8 planes per neighbour, staged through the device, NCCL send/recv)",
           "pipeline": f"{slabs} z-slabs per rank", "checksum_f": float(chk.item())}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None] + list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=40)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
