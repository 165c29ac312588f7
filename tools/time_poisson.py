"""Times the pressure projection on one B200: update! + myproject! on a two-phase problem (sphere, density ratio
1/λρ), random divergent velocity.  Prints one JSON line per grid: iterations, ms per iteration, algorithmic GB/s at 17 s B per cell
and iteration (ifadv_poisson.cuh header)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import interfaceadvection.jl_b200 as ia  # noqa: E402


def run(N, T, itmx, perdir=()):
    n0 = N[0]
    sim = ia.TwoPhaseSimulation(N, (0,) * len(N), float(n0), T=T, lam_rho=1e-3, dt=0.25, psolver="Poisson", perdir=perdir,
                                InterfaceSDF=lambda x: ((x - n0 / 2) ** 2).sum(-1).sqrt() - n0 / 4)
    a, c, b = sim.flow, sim.intf, sim.pois
    gen = torch.Generator(device="cuda").manual_seed(5)
    a.u.copy_(0.1 * torch.randn(a.u.shape, generator=gen, device="cuda", dtype=T))
    ia.BC(a.u, a.uBC, False, perdir)
    ia.updateL(a.mu0, c.f, c.lam_rho, perdir, fill_one=True)
    ia.update(b)
    ctx = ia.context_for(c.f)
    # fixed number of iterations through psolver (the source is ∇·u)
    u0 = a.u.clone()
    res = []
    for rep in range(1 if len(sys.argv) > 1 else 3):
        a.u.copy_(u0); a.p.zero_()
        ia.myproject(a, b, 1.0) if rep == 0 and itmx is None else None
        b.z.zero_(); b.x.zero_()
        D = len(N)
        for i in range(D):
            hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(D)) + (i,)
            lo = (slice(1, -1),) * D + (i,)
            b.z[(slice(1, -1),) * D] += a.u[hi] - a.u[lo]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record()
        n = ia.psolver(b, itmx=itmx or 6000)
        e1.record(); torch.cuda.synchronize()
        res.append((n, e0.elapsed_time(e1), ctx.launches - l0))
    n, ms, launches = min(res, key=lambda r: r[1])
    cells = int(np.prod(N))
    s = torch.tensor([], dtype=T).element_size()
    print(json.dumps({"what": "psolver", "N": list(N), "dtype": str(T).split(".")[-1], "perdir": list(perdir), "iterations": n,
                      "ms": round(ms, 3), "ms_per_iteration": round(ms / max(n, 1), 4), "r2": b.r2[-1], "launches": launches,
                      "algorithmic_GBps": round(17 * s * cells * n / (ms * 1e-3) / 1e9, 1),
                      "Gcell_iterations_per_s": round(cells * n / (ms * 1e-3) / 1e9, 2)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:  # N dtype itmx: one configuration (for ncu)
        n = int(sys.argv[1])
        run((n, n, n), getattr(torch, sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    run((64, 64, 64), torch.float32, 200)
    run((256, 256, 256), torch.float32, 200)
    run((256, 256, 256), torch.float64, 200)
    run((512, 512, 512), torch.float32, 100)
    run((512, 512, 512), torch.float32, 100, perdir=(1, 2))
