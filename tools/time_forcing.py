"""Times the explicit-forcing kernels on cuda:0 with CUDA events: viscSurfTenρu! (visc only / surface tension only /
both), updateU!, updateL!, and the whole MPFMomStep! with forcing next to the transport-only step.  C4: 512^3 Float32 bubble (default) or
`--n 256 --f64`.  Prints one JSON line.  bench.py imports measure() for its `forcing` sub-line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(n=512, f64=False, reps=10, dev="cuda"):
    import numpy as np
    import torch
    import interfaceadvection.jl_b200 as ia
    from interfaceadvection.jl_b200 import configs

    N = (n,) * 3
    T = torch.float64 if f64 else torch.float32
    es = 8 if f64 else 4
    R = n / 8
    cen = torch.tensor([n / 2, n / 2, n / 4], device=dev)
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), float(n), T=T, lam_mu=1e-2, lam_rho=1e-3, eta=0.05, nu=0.01, g=(0.0, 0.0, -0.001),
                                InterfaceSDF=lambda x: R - ((x - cen.to(x.dtype)) ** 2).sum(-1).sqrt(), perdir=(1, 2), dt=0.25, device=dev)
    a, c = sim.flow, sim.intf
    case = configs.make_case(N, dtype="float64" if f64 else "float32", kind="C4", vel="enright")
    a.u.copy_(ia.from_numpy(np.asfortranarray(case["u"] * 0.5), device=dev))
    del case
    ia.BC(a.u, a.uBC, False, a.perdir)
    a.dt[:] = [0.25]
    cells = float(np.prod(N))

    def t(fn, k=reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    out = {"grid": list(N), "dtype": "f64" if f64 else "f32", "reps": reps,
           "what": "explicit forcing of MPFMomStep! (SURVEY 8f row 1): viscSurfTenrhou! (mu, eta), updateU! (gravity), updateL!; bubble R = n/8, "
                   "mu = 0.01, eta = 0.05, lambda_mu = 1e-2, lambda_rho = 1e-3, periodic x,y"}
    vs = lambda mu, eta: ia.viscSurfTenrhou(a.f, a.u, a.sigma, c.f, c.alpha, c.nhat, c.ff, c.lam_mu, mu, c.lam_rho, eta, a.perdir)
    out["visc_ms"] = t(lambda: vs(0.01, None))
    out["surften_ms"] = t(lambda: vs(None, 0.05))
    out["visc_surften_ms"] = t(lambda: vs(0.01, 0.05))
    out["update_u_ms"] = t(lambda: ia.updateU(a.u, c.rhou, c.nhat, a.f, 1e-6, c.f, c.lam_rho, 0.0, None, a.uBC, 1.0))
    out["update_l_ms"] = t(lambda: ia.updateL(a.mu0, c.f, c.lam_rho, a.perdir, fill_one=True))
    # algorithmic bytes per cell: visc reads u(3)+f, writes r(3); updateU reads ρu,ρu⁰,forcing(9)+f, writes ρu,u,forcing(9); updateL reads f, writes μ₀(3)
    out["algorithmic_bytes_per_cell"] = {"visc": 7 * es, "update_u": 19 * es, "update_l": 4 * es}
    out["visc_GBs"] = cells * 7 * es / out["visc_ms"] / 1e6
    out["update_u_GBs"] = cells * 19 * es / out["update_u_ms"] / 1e6
    out["update_l_GBs"] = cells * 4 * es / out["update_l_ms"] / 1e6
    ia.BC(a.u, a.uBC, False, a.perdir)
    out["step_transport_only_ms"] = t(lambda: ia.mom_advect_step(a, c, 0.25), k=5)
    out["step_with_forcing_ms"] = t(lambda: ia.mom_step_forcing(a, c, 0.25), k=5)
    out["finite"] = bool(torch.isfinite(a.u).all().item() and torch.isfinite(c.f).all().item())
    del sim, a, c
    ia.api._contexts.clear()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    print(json.dumps(measure(args.n, args.f64, args.reps)))
