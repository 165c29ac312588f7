#!/usr/bin/env python
"""bench.py -- VOF+CMOM advection throughput of the B200-native path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one CMOM advection step = the transport half of MPFMomStep! (src/flow.jl:61,69-70,74,89-92):
2 x (u2ρu! + BC! + advectfq!) + the midpoint f⁰ = 2·D fused directional sweeps.  MPCFL and the
Poisson solve are excluded (timed separately by the reference's users).

Workload at N=1: BASELINE.json config 4, the 512³ Float32 rising-bubble grid (periodic x,y; λρ=1e-3; Koren; WH normals) --
the grid north_star quotes its roofline target on -- with a THREE-COMPONENT discretely solenoidal velocity (the discrete-curl
LeVeque/Enright field), so every sweep direction carries a non-zero flux; the Taylor-Green line of round 1 (w ≡ 0) is measured
beside it ("tgv_line").  N>1 (torchrun, one rank per GPU): the same 512x512x512 block per GPU, stacked along z (weak scaling;
z-slabs, per-sweep NCCL ghost-plane exchange inside the library), preceded by a small-grid check that the N-GPU result equals the
1-GPU result bit for bit.  Every line also carries "c5": BASELINE config 5's named slab, 2048x1024x128 per GPU, at the same N.
`--workload C5_strong_2048x1024x512_f32` splits that fixed grid over the ranks (strong scaling).

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


def _wl(N, dtype, perdir, kind, vel=None, strong=False):
    return dict(N=tuple(N), dtype=dtype, perdir=tuple(perdir), kind=kind, vel=vel, strong=strong)


WORKLOADS = {
    # N per GPU (weak) or the global grid (strong=True, split along z), dtype, perdir, SDF kind, velocity generator
    "C4_bubble_512_f32": _wl((512, 512, 512), "float32", (1, 2), "C4", "enright"),
    "C4_bubble_512_f32_tgv": _wl((512, 512, 512), "float32", (1, 2), "C4"),  # round-1 headline: Taylor-Green, w ≡ 0
    "C3_dambreak_512x256x256_f32": _wl((512, 256, 256), "float32", (), "C3", "enright"),
    "C4_bubble_256_f32": _wl((256, 256, 256), "float32", (1, 2), "C4", "enright"),
    "C4_bubble_256_f64": _wl((256, 256, 256), "float64", (1, 2), "C4", "enright"),
    "C4_bubble_128_f32": _wl((128, 128, 128), "float32", (1, 2), "C4", "enright"),
    "C4_bubble_64_f32": _wl((64, 64, 64), "float32", (1, 2), "C4", "enright"),
    "C5_sloshing_2048x1024x128_f32": _wl((2048, 1024, 128), "float32", (), "C5", "enright"),
    "C5_strong_2048x1024x512_f32": _wl((2048, 1024, 512), "float32", (), "C5", "enright", strong=True),
    # pure-VOF configs (one step = advect! = D directional sweeps): reported for completeness, single GPU, no e2e / CPU legs
    "C2_enright_256_f32": _wl((256, 256, 256), "float32", (), "C2"),
    "C2_enright_256_f64": _wl((256, 256, 256), "float64", (), "C2"),
    "C1_zalesak_128_f64": _wl((128, 128), "float64", (), "C1"),
}
VOF_KINDS = ("C1", "C2")
DEFAULT = "C4_bubble_512_f32"
C5 = "C5_sloshing_2048x1024x128_f32"
CPU_SAMPLE_N = (256, 256, 256)  # bounded sample of the same workload for the cpu_baseline leg of the b200 arm


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
def cpu_case(w, N=None):
    from interfaceadvection.jl_b200 import configs
    from oracle import pyoracle as O

    N = tuple(N or w["N"])
    dtype, perdir, kind = w["dtype"], w["perdir"], w["kind"]
    T = np.dtype(dtype).type
    case = configs.make_case(N, dtype=dtype, kind=kind, vel=w["vel"])
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


def time_cpu(w, steps, warmup, N=None):
    from oracle import pyoracle as O

    O.build()
    # state the thread count explicitly: torchrun exports OMP_NUM_THREADS=1, which would pin the OpenMP build to one core
    O.set_num_threads(os.cpu_count() or 1)
    c = cpu_case(w, N)
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
    """The reference arm: the reference's own CPU implementation of the path -- here its C++/OpenMP restatement (the package is
    100 % Julia and no Julia exists in this image) -- on the box's host cores, ON THE WORKLOAD'S OWN GRID.  Each timed step is one
    full CMOM advection step of that grid; at 512³ that is several seconds, so the number of timed steps is bounded (<= 3)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload or DEFAULT
    w = WORKLOADS[wl]
    N, dtype, perdir = w["N"], w["dtype"], w["perdir"]
    cells = math.prod(N)
    steps = max(1, min(args.steps, 3 if cells > 64e6 else (10 if cells > 8e6 else 40)))
    warm = 1
    val, ms, cores, c = time_cpu(w, steps, warm)
    sample = (f"{wl}: the full {'x'.join(map(str, N))} {dtype} grid, {steps} timed step(s) after {warm} warm-up (bounded: one step is "
              f"{ms / 1e3:.1f} s of CPU work); restated reference algorithm (C++/OpenMP, un-fused pass structure), not the Julia package; "
              f"{cores} threads on {cpu_model()} ({os.cpu_count()} logical cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "steps_requested": args.steps, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": wl, "grid_per_gpu": list(N), "perdir": list(perdir), "limiter": "Koren", "normal_scheme": "WH",
                   "lambda_rho": 1e-3, "velocity": w["vel"] or "config default", "measured_on": "the workload's own grid (host cores)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def fill_state_chunked(ia, torch, configs, sim, N, w, dev, nzc=64):
    """Initial f and u of a large single-GPU grid, sampled from the analytic generators in z-chunks so that the temporaries of the
    generators (a dozen full-size fields) stay small next to the 26-field working set."""
    D = len(N)
    T = sim.intf.f.dtype
    if D != 3 or N[2] <= nzc:
        case = configs.make_case(N, dtype=w["dtype"], device=dev, kind=w["kind"], vel=w["vel"])
        ia.applyVOF(sim.intf.f, sim.intf.alpha, sim.intf.nhat, case["sdf"])
        sim.flow.u.copy_(case["u"])
    else:
        for z0 in range(0, N[2], nzc):
            nl = min(nzc, N[2] - z0)
            Nl = (N[0], N[1], nl)
            case = configs.make_case(N, dtype=w["dtype"], device=dev, kind=w["kind"], vel=w["vel"], Nl=Nl, origin=(0, 0, z0))
            Ngl = tuple(n + 2 for n in Nl)
            f = ia.jl_zeros(Ngl, T, dev); al = ia.jl_zeros(Ngl, T, dev); nh = ia.jl_zeros(Ngl + (3,), T, dev)
            ia.applyVOF(f, al, nh, case["sdf"], origin=(0, 0, z0))
            sim.intf.f[:, :, 1 + z0:1 + z0 + nl] = f[:, :, 1:-1]
            sim.flow.u[:, :, 1 + z0:1 + z0 + nl, :] = case["u"][:, :, 1:-1, :]
            del f, al, nh, case
        torch.cuda.empty_cache()
    ia.BCf(sim.intf.f, w["perdir"])
    sim.intf.f0.copy_(sim.intf.f)
    ia.BC(sim.flow.u, (0,) * D, False, w["perdir"])
    sim.flow.u0.copy_(sim.flow.u)


class _Single:
    """One GPU: the exported simulation surface (TwoPhaseSimulation + the transport half of MPFMomStep!)."""

    def __init__(self, ia, torch, configs, w, dev):
        N, D = w["N"], len(w["N"])
        T = getattr(torch, w["dtype"])
        self.ia, self.vof = ia, w["kind"] in VOF_KINDS
        self.sim = ia.TwoPhaseSimulation(N, (0,) * D, float(N[0]), T=T, lam_rho=1e-3, perdir=w["perdir"], U=1.0, dt=1.0, device=dev)
        fill_state_chunked(ia, torch, configs, self.sim, N, w, dev)
        self.contexts = [ia.context_for(self.sim.intf.f)]

    def step(self):
        if self.vof:
            # advect!(a,c): pure VOF with u⁰, u (advection.jl:17-23); without the optional ρuf output, as the 4s+1 B/cell accounting assumes
            self.ia.advect(self.sim.flow, self.sim.intf, check=False, want_rhouf=False)
        else:
            self.ia.mom_advect_step(self.sim.flow, self.sim.intf, 1.0)
        self.sim.flow.dt.append(1.0)  # fixed Δt; advances the sweep-order rotation like push!(Δt) would

    def mass(self):
        return self.ia.sum_inside(self.sim.intf.f)


def make_runner(ia, torch, w, rank, world, dev):
    from interfaceadvection.jl_b200 import configs

    if world == 1:
        return _Single(ia, torch, configs, w, dev), w["N"]
    from interfaceadvection.jl_b200 import slab
    N = w["N"]
    if w["strong"]:
        assert N[2] % world == 0, "the strong-scaling grid must split evenly"
        N = (N[0], N[1], N[2] // world)
    return slab.SlabRunner(N, w["dtype"], w["perdir"], w["kind"], rank, world, dev, vel=w["vel"]), N


def timed_run(torch, dist, runner, steps, warmup, dev, rank, local, clocks=True):
    """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream, max over ranks."""
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    m0 = runner.mass()
    for _ in range(warmup):
        runner.step()
    barrier()
    l0 = sum(c.launches for c in runner.contexts)
    for c in runner.contexts:
        c.profile(True)
    sampler = ClockSampler(local) if (clocks and rank == 0) else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        runner.step()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    ck = sampler.stop(t0, t1) if sampler else None
    kms, kn, fms, fn_, per_dir = 0.0, 0, 0.0, 0, {}
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
    return dict(ms=ms, kms=kms, kn=kn, fms=fms, fn=fn_, per_dir=per_dir, launches=int(launches), clocks=ck,
                mass_drift_rel=abs(m1 - m0) / abs(m0) if m0 else None)


def hbm_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    return peak, src


def roofline_of(w, wl, N_gpu, r, steps):
    D = len(N_gpu)
    s = 4 if w["dtype"] == "float32" else 8
    vof = w["kind"] in VOF_KINDS
    cells_gpu = math.prod(N_gpu)
    peak, peak_src = hbm_peak()
    bytes_launch = (vof_bytes_per_cell_sweep(D, s) if vof else algorithmic_bytes_per_cell_sweep(D, s)) * cells_gpu
    avg_ms = r["kms"] / max(r["kn"], 1)
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9 if r["kn"] else None
    # fused first sweeps (u2rhou + BC folded in): 3 fewer input streams -> (2D+4)s+1 B per cell
    fbytes = ((2 * D + 4) * s + 1) * cells_gpu
    favg = r["fms"] / max(r["fn"], 1)
    fused_info = {"avg_launch_ms": favg, "launches_timed": r["fn"], "algorithmic_bytes_per_launch": fbytes,
                  "achieved_gbs": (fbytes / (favg * 1e-3) / 1e9)} if r["fn"] else None
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(wl, {}).get("dram_bytes_per_launch")
        traffic_src = tr.get(wl, {}).get("source")
    except Exception:
        pass
    ms_step = r["ms"] / steps
    step_bytes = (vof_bytes_per_cell_step(D, s) if vof else algorithmic_bytes_per_cell_step(D, s)) * cells_gpu
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
           "traffic": traffic,
           "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture "
                            f"({traffic_src or 'profiles/'}), NOT measured in this run") if traffic else None,
           "kernel": "ifadv::xrow_kernel (x) / ifadv::along2_kernel (y, z): fused VOF+CMOM directional sweep, standard 13s+1 B/cell form, "
                     "average over the three directions",
           "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_launch,
           # both advectfq! calls of MPFMomStep! pass ONE velocity array for u¹ and u²: the kernels that run (SAMEU) read one stream less
           "executed_bytes_per_launch_single_velocity_array": ((2 * D + 6) * s + 1) * cells_gpu if not vof else None,
           "frac_on_executed_bytes": ((((2 * D + 6) * s + 1) * cells_gpu / (avg_ms * 1e-3) / 1e9) / peak) if (r["kn"] and not vof) else None,
           "avg_launch_ms": avg_ms, "launches_timed": r["kn"],
           "sweep_share_of_step": ((r["kms"] + r["fms"]) / r["ms"]) if r["ms"] else None, "fused_first_sweep": fused_info,
           "ms_per_launch_by_direction": r["per_dir"],
           "step_frac_of_roofline": (step_bytes / (ms_step * 1e-3) / 1e9) / peak}
    if vof:
        out["kernel"] = "fused pure-VOF directional sweep (ifadv::along2_kernel y/z, ifadv::march_kernel x; 2-D: ifadv::sweep_kernel), 4s+1 B/cell"
        out["fused_first_sweep"] = None
    return out


def slab_bitwise_check(ia, torch, dist, rank, world, dev):
    """Before timing at N > 1: a small global grid advanced by the N-rank slab decomposition and, on every rank, by the single-GPU
    path (the oracle of the multi-GPU path).  Float64 (IEEE-exact arithmetic): the owned cells of f and ρu must be
    BIT-IDENTICAL.  Float32 is built with FMA contraction, and the boundary / interior instantiations of a kernel contract
    differently, so a plane that is "interior" on one GPU and "next to a slab end" on N GPUs may differ in the last bit: reported as
    the largest absolute difference (a few 1e-8 on O(1) fields), required to stay below 1e-6."""
    from interfaceadvection.jl_b200 import configs, slab

    N = (96, 64, 24)
    perdir = (1, 2)
    Ng = (N[0], N[1], N[2] * world)
    out = {"grid_per_gpu": list(N), "steps": 2, "velocity": "enright (three components)", "interface": "sphere across the rank 0 / rank 1 slab boundary"}
    for dtype in ("float64", "float32"):
        T = getattr(torch, dtype)
        case = configs.make_case(Ng, dtype=dtype, device=dev, kind="C4", vel="enright")
        sdf = configs.sdf_sphere([N[0] / 2, N[1] / 2, N[2] * 1.0], N[2] * 0.6, inside_dark=False)
        sim = ia.TwoPhaseSimulation(Ng, (0, 0, 0), float(N[0]), T=T, lam_rho=1e-3, InterfaceSDF=sdf, perdir=perdir, U=1.0, dt=1.0, device=dev)
        sim.flow.u.copy_(case["u"]); ia.BC(sim.flow.u, (0, 0, 0), False, perdir)
        g = slab.SlabGeom(rank, world, N[2], slab.G_DEFAULT, False)
        nzg = N[2] * world
        zidx = torch.tensor([min(max(g.z_origin + l, 0), nzg + 1) for l in range(g.nz_local + 2)], device=dev)
        run = slab.SlabRunner(N, dtype, perdir, "C4", rank, world, dev, fields=(sim.intf.f.index_select(2, zidx), sim.flow.u.index_select(2, zidx)))
        for _ in range(2):
            run.step()
            ia.mom_advect_step(sim.flow, sim.intf, 1.0); sim.flow.dt.append(1.0)
        torch.cuda.synchronize()
        sl = slice(1 + rank * N[2], 1 + (rank + 1) * N[2])
        df = (sim.intf.f[1:-1, 1:-1, sl] - run.owned_f()).abs().max()
        dr = (sim.intf.rhou[1:-1, 1:-1, sl, :] - run.owned_rhou()).abs().max()
        ok = torch.equal(sim.intf.f[1:-1, 1:-1, sl], run.owned_f()) and torch.equal(sim.intf.rhou[1:-1, 1:-1, sl, :], run.owned_rhou())
        v = torch.tensor([1.0 if ok else 0.0, -float(df), -float(dr)], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MIN)
        key = "f64" if dtype == "float64" else "f32"
        out[key + "_owned_cells_bit_identical_to_1gpu"] = bool(v[0].item() == 1.0)
        out[key + "_max_abs_diff"] = {"f": -float(v[1].item()), "rhou": -float(v[2].item())}
        del run, sim
        ia.api._contexts.clear()
    out["pass"] = bool(out["f64_owned_cells_bit_identical_to_1gpu"] and max(out["f32_max_abs_diff"].values()) <= 1e-6)
    return out


def measure_workload(ia, torch, dist, wl, rank, world, dev, local, steps, warmup, clocks):
    w = WORKLOADS[wl]
    runner, N_gpu = make_runner(ia, torch, w, rank, world, dev)
    r = timed_run(torch, dist, runner, steps, warmup, dev, rank, local, clocks=clocks)
    cells = math.prod(N_gpu) * world
    value = cells * steps / (r["ms"] * 1e-3) / 1e9
    roof = roofline_of(w, wl, N_gpu, r, steps)
    sent = getattr(runner, "bytes_sent", 0)
    w = dict(w, _transport=getattr(runner, "transport", "none"))
    del runner
    if dist is not None:
        dist.barrier()  # every rank has dropped its mappings of the neighbours' buffers before anybody frees them
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    ia.api._contexts.clear()
    return w, N_gpu, r, value, roof, sent


def run_b200(args):
    import torch

    import interfaceadvection.jl_b200 as ia

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
    wl = args.workload or DEFAULT
    w = WORKLOADS[wl]
    D = len(w["N"])
    vof = w["kind"] in VOF_KINDS
    if vof:
        if world > 1:
            raise SystemExit("the pure-VOF workloads are single-GPU report lines")
        args.no_e2e = args.no_cpu = args.no_extra = True
    if wl != DEFAULT:
        args.no_extra = True

    # ---- N > 1: the N-GPU result must equal the 1-GPU result before anything is timed ----
    bitwise = slab_bitwise_check(ia, torch, dist, rank, world, dev) if world > 1 else None
    ia.api._contexts.clear()

    # ---- main line ----
    w, N_gpu, r, value, roofline, sent = measure_workload(ia, torch, dist, wl, rank, world, dev, local, args.steps, args.warmup, True)
    ms = r["ms"]

    # ---- beside it: the Taylor-Green line of round 1 (w ≡ 0) and BASELINE config 5's named slab at the same N ----
    tgv_line, c5_line = None, None
    if not args.no_extra:
        k = min(args.steps, 10)
        _, _, r2, v2, roof2, _ = measure_workload(ia, torch, dist, "C4_bubble_512_f32_tgv", rank, world, dev, local, k, 3, False)
        tgv_line = {"workload": "C4_bubble_512_f32_tgv", "velocity": "Taylor-Green (u_z ≡ 0: the z-sweep carries no flux)", "value": v2,
                    "ms_per_step": r2["ms"] / k, "steps": k, "ms_per_launch_by_direction": r2["per_dir"],
                    "step_frac_of_roofline": roof2["step_frac_of_roofline"],
                    "delta_ms_per_launch_vs_main": {d: r["per_dir"].get(d, 0.0) - r2["per_dir"].get(d, 0.0) for d in r2["per_dir"]}}
        _, N5, r5, v5, roof5, sent5 = measure_workload(ia, torch, dist, C5, rank, world, dev, local, k, 3, False)
        c5_line = {"workload": C5, "grid_per_gpu": list(N5), "global_grid": [N5[0], N5[1], N5[2] * world], "scaling": "weak", "value": v5,
                   "unit": UNIT, "ms_per_step": r5["ms"] / k, "steps": k, "step_frac_of_roofline": roof5["step_frac_of_roofline"],
                   "ms_per_launch_by_direction": r5["per_dir"], "sweep_share_of_step": roof5["sweep_share_of_step"],
                   "nccl_bytes_sent_per_rank_per_step": int(sent5 / (k + 3)) if world > 1 else 0,
                   "note": "BASELINE config 5 (sloshing tank 2048x1024x1024 over 8 GPUs): this is its per-GPU slab at N GPUs; weak-scaling "
                           "efficiency of the named shape = value(N) / (N * value(1)) of this sub-line"}

    # ---- next to the path: the explicit forcing of MPFMomStep! on the same grid, timed on its own ----
    forcing_line = None
    if not args.no_extra and world == 1:
        import importlib.util
        spec = importlib.util.spec_from_file_location("time_forcing", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "time_forcing.py"))
        tf = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tf)
        ia.api._contexts.clear()
        torch.cuda.empty_cache()
        forcing_line = tf.measure(w["N"][0], w["dtype"] == "float64", 5, dev)

    # ---- next to the path: the reference's Jacobi-PCG pressure solver on the same grid, timed on its own ----
    projection_line = None
    if not args.no_extra and world == 1 and D == 3:
        import importlib.util
        spec = importlib.util.spec_from_file_location("time_poisson", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "time_poisson.py"))
        tp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tp)
        ia.api._contexts.clear()
        torch.cuda.empty_cache()
        projection_line = tp.measure(w["N"][0], w["dtype"] == "float64", 50, dev, cpu_n=0 if args.no_cpu else 128)
        ia.api._contexts.clear()
        torch.cuda.empty_cache()
        try:  # WaterLily's default psolver (MultiLevelPoisson, inproject!'s second method flow.jl:343-347) on the same problem
            spec = importlib.util.spec_from_file_location("time_mlpoisson", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "time_mlpoisson.py"))
            tm = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(tm)
            projection_line["multigrid"] = tm.measure(w["N"][0], w["dtype"] == "float64", 4, dev)
        except ia.IfadvError as e:  # grid without three multigrid levels (WaterLily's own constructor assertion)
            projection_line["multigrid"] = {"unavailable": str(e)}
        ia.api._contexts.clear()
        torch.cuda.empty_cache()

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D/D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        if world == 1:
            e2e = run_e2e(ia, torch, w, dev, min(args.steps, args.e2e_steps))
        else:
            e2e = run_e2e_slabs(ia, torch, dist, w, N_gpu, dev, rank, world, min(args.steps, args.e2e_steps))

    # ---- CPU baseline (rank 0, N=1 only): a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        Ns = tuple(min(a, b) for a, b in zip(w["N"], CPU_SAMPLE_N)) if D == 3 else w["N"]
        v, cms, cores, c = time_cpu(w, args.cpu_steps, 1, N=Ns)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{'x'.join(map(str, c['N']))} {w['dtype']} box of the same workload (same SDF / velocity generators scaled to the box), "
                         f"{args.cpu_steps} steps after 1 warm-up, {cms:.0f} ms/step; restated reference algorithm (C++/OpenMP, un-fused pass "
                         f"structure), not the Julia package; {cpu_model()}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if w["strong"] else "weak", "vs_baseline": None,
            "dtype": "f32" if w["dtype"] == "float32" else "f64", "data": "synthetic",
            "config": {"workload": wl, "grid_per_gpu": list(N_gpu), "global_grid": [N_gpu[0], N_gpu[1], N_gpu[2] * world] if D == 3 else list(N_gpu),
                       "perdir": list(w["perdir"]), "limiter": "Koren", "normal_scheme": "WH", "lambda_rho": 1e-3,
                       "velocity": {"enright": "discrete curl of the LeVeque/Enright vector potential: three components, exactly discretely "
                                               "solenoidal", None: "config default (Taylor-Green, u_z ≡ 0 in 3-D; C1: rigid rotation; C2: Enright)"}[w["vel"]],
                       "step": "pure-VOF advection step = advect! = D fused sweeps" if vof else "CMOM advection step = 2 x (u2rhou + BC + advectfq) + midpoint f0 = 6 fused sweeps (u2rhou+BC folded into the first sweep of each group); the u0<-u / f0<-f copies and the midpoint run on a second stream underneath the sweeps",
                       "l2": "working set >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "parallelism": (f"z-slab x{world}: 3 ghost planes per neighbour side, exchange of f (3 planes up / 2 down), rho-u (2 / 2) and "
                                       f"c-bar after every directional sweep; transport: {w.get('_transport')}") if world > 1 else "single GPU",
                       "nccl_bytes_sent_per_rank_per_step": int(sent / (args.steps + args.warmup)) if world > 1 else 0,
                       "mass_drift_rel": r["mass_drift_rel"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": r["launches"], "clocks": r["clocks"],
            "slab_check": bitwise, "tgv_line": tgv_line, "c5": c5_line, "forcing": forcing_line, "projection": projection_line,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _host_state(ia, torch, configs, w, Nl, origin, dev, perdir):
    """f, u of a (sub-)box as pinned host buffers holding the column-major bytes."""
    D = len(Nl)
    T = getattr(torch, w["dtype"])
    Ng = tuple(n + 2 for n in Nl)
    Nglob = w["_Nglobal"]
    case = configs.make_case(Nglob, dtype=w["dtype"], device=dev, kind=w["kind"], vel=w["vel"], Nl=Nl if origin else None, origin=origin)
    f = ia.jl_zeros(Ng, T, dev); al = ia.jl_zeros(Ng, T, dev); nh = ia.jl_zeros(Ng + (D,), T, dev)
    ia.applyVOF(f, al, nh, case["sdf"], origin=origin); ia.BCf(f, perdir)
    u = case["u"]; ia.BC(u, (0,) * D, False, perdir)
    fh = torch.empty(tuple(reversed(Ng)), dtype=T, pin_memory=True)
    uh = torch.empty((D,) + tuple(reversed(Ng)), dtype=T, pin_memory=True)
    rh = torch.empty((D,) + tuple(reversed(Ng)), dtype=T, pin_memory=True)
    fh.copy_(f.permute(*reversed(range(D)))); uh.copy_(u.permute(*reversed(range(D + 1))))
    return fh, uh, rh, f


def run_e2e(ia, torch, w, dev, steps):
    """Same metric through ifadv_mom_advect_step_host: pinned host f,u -> device, one CMOM step, f,ρu -> host."""
    from interfaceadvection.jl_b200 import configs

    N, perdir, dtype = w["N"], w["perdir"], w["dtype"]
    D = len(N)
    Ng = tuple(n + 2 for n in N)
    w = dict(w, _Nglobal=N)
    fh, uh, rh, f = _host_state(ia, torch, configs, w, N, None, dev, perdir)
    del f
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


def run_e2e_slabs(ia, torch, dist, w, N_gpu, dev, rank, world, steps):
    """The e2e leg at N > 1 GPUs: every rank keeps ITS z-slab of the global state (owned planes + 8 overlap planes per interior
    end) in pinned host memory and advances it with ifadv_mom_advect_step_host; after each step the overlap planes of the host
    f are refreshed from the neighbours' owned planes (staged through the device, NCCL send/recv) -- the host-level halo
    exchange a multi-process caller has to do.  (The host entry runs whole steps on a private device copy, so it uses the wide
    overlap of 8 planes -- the reach of one full step -- instead of the per-sweep exchange of the device-resident path.)
    Timed region: barrier, steps x (host step + exchange), barrier; max over ranks."""
    from interfaceadvection.jl_b200 import configs, slab

    W = 8
    g = slab.SlabGeom(rank, world, N_gpu[2], W, 3 in w["perdir"])
    Nl = (N_gpu[0], N_gpu[1], g.nz_local)
    lperdir = g.local_perdir(w["perdir"])
    w = dict(w, _Nglobal=(N_gpu[0], N_gpu[1], N_gpu[2] * world))
    # every rank pins 7 fields of its slab (~4 GB at 512^3): if one rank cannot, all ranks skip the leg together
    ok = torch.ones(1, device=dev)
    fh = uh = rh = fdev = None
    try:
        fh, uh, rh, fdev = _host_state(ia, torch, configs, w, Nl, (0, 0, g.z_origin), dev, lperdir)
    except Exception as e:  # noqa: BLE001
        print(f"[bench] rank {rank}: cannot build / pin the host buffers of the e2e leg: {e}", file=sys.stderr)
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        return None
    Ngl = tuple(fdev.shape)
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    ctx = ia.Context(Ngl, w["dtype"], dev.index or 0)
    lim, ns = ia.LIMITERS["Koren"], ia.NORMAL_SCHEMES["WH"]
    o = g.owned
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
    out = {"value": math.prod(N_gpu) * world * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
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
    ap.add_argument("--no-extra", action="store_true", help="skip the tgv_line / c5 sub-lines of the default workload")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=6)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
