"""Shared builders for the parity tests: a case -> the full reference array set as Fortran numpy arrays."""
import numpy as np

from oracle import pyoracle as O
from interfaceadvection.jl_b200 import configs


def make_state(N, kind, dtype, perdir=None, uBC=None, scale_u=1.0, shift=None):
    """Build (f, u) like cVOF/Flow construction would: f from applyVOF!+BCf!, u with BC! applied."""
    case = configs.make_case(tuple(N), dtype=np.dtype(dtype).name, kind=kind)
    perdir = case["perdir"] if perdir is None else tuple(perdir)
    D = len(N)
    Ng = tuple(n + 2 for n in N)
    f = O.zeros(Ng, dtype); al = O.zeros(Ng, dtype); nh = O.zeros(Ng + (D,), dtype)
    sdf = case["sdf"]
    if shift is not None:
        base = sdf
        sdf = lambda x: base(x - np.asarray(shift, dtype=x.dtype))
    O.applyVOF(f, al, nh, sdf)
    O.BCf(f, perdir)
    u = np.asfortranarray((case["u"] * dtype(scale_u)).astype(dtype))
    uBC = tuple([0.0] * D) if uBC is None else tuple(uBC)
    if any(uBC):
        for i in range(D):
            u[..., i] += dtype(uBC[i])
    O.BC(u, uBC, False, perdir)
    return dict(N=tuple(N), D=D, Ng=Ng, dtype=dtype, perdir=perdir, uBC=uBC, f=f, u=u, lam_rho=case["lam_rho"])


def alloc_cmom(st):
    """The array set of advectfq! (flow.jl:157-160) with the reference's aliasing uStar≡n̂, dilaU≡α."""
    Ng, D, T = st["Ng"], st["D"], st["dtype"]
    z = lambda *s: O.zeros(s, T)
    a = dict(ff=z(*Ng), alpha=z(*Ng), nhat=z(*Ng, D), cbar=np.zeros(Ng, dtype=np.int8, order="F"), rhou=z(*Ng, D), r=z(*Ng, D),
             Phi=z(*Ng), rhouf=z(*Ng, D), drho=O.zeros(Ng + (D,), T))
    a["drho"][...] = 1
    return a


def oracle_cmom_call(st, f, u1, u2, uOld, rhou, dt, dirO, lam="Koren", scheme="WH", arrays=None):
    a = alloc_cmom(st) if arrays is None else arrays
    a["rhou"][...] = rhou
    status, rep = O.advectVOFrhouu(f, a["ff"], a["alpha"], a["nhat"], u1, u2, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"],
                                   a["nhat"], uOld, a["alpha"], a["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False,
                                   dirO)
    return status, rep, a


def oracle_mom_advect_step(st, f, u, dt, dirO, lam="Koren", scheme="WH", omp=False):
    """Transport part of MPFMomStep! with prescribed velocity (flow.jl:61,69-70,74,89-92) on the oracle."""
    T = st["dtype"]
    a = alloc_cmom(st)
    u0 = u.copy(order="F"); f0 = f.copy(order="F")
    O.u2rhou(a["rhou"], u0, f0, st["lam_rho"]); O.BC(a["rhou"], st["uBC"], False, st["perdir"])
    O.advectVOFrhouu(f0, a["ff"], a["alpha"], a["nhat"], u0, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u,
                     a["alpha"], a["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO, omp=omp)
    f0[...] = (f0 + f) * T(0.5)
    f0[...] = f
    O.u2rhou(a["rhou"], u0, f, st["lam_rho"]); O.BC(a["rhou"], st["uBC"], False, st["perdir"])
    O.advectVOFrhouu(f, a["ff"], a["alpha"], a["nhat"], u, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u0,
                     a["alpha"], a["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO, omp=omp)
    return a["rhou"]


def inside(a, D):
    sl = tuple([slice(1, -1)] * D)
    return a[sl]


def dirO_for(nsteps_done, D):
    n = 1 + nsteps_done  # length(Δt)
    return tuple((n + i) % D + 1 for i in range(1, D + 1))


def second_velocity(st, seed, scale, amp):
    """A velocity field that differs from st["u"] in VALUE: scale·u + a smooth, non-solenoidal perturbation (so the dilation
    terms see ∂u¹ ≠ ∂u²), periodic with the box, with the case's BC! applied."""
    T, D, Ng = st["dtype"], st["D"], st["Ng"]
    idx = np.indices(Ng).astype(np.float64)
    ph = [2 * np.pi * idx[k] / (Ng[k] - 2) for k in range(D)]
    v = np.empty(Ng + (D,), dtype=T, order="F")
    for i in range(D):
        w = np.sin(ph[0] + 0.37 * seed + 0.9 * i) * np.cos(ph[1] - 0.21 * seed + 0.4 * i)
        if D == 3:
            w = w * np.cos(ph[2] + 0.3 * seed - 0.7 * i)
        v[..., i] = (T(scale) * st["u"][..., i].astype(np.float64) + amp * w).astype(T)
    O.BC(v, st["uBC"], False, st["perdir"])
    return v


def oracle_mom_step_hook(st, f, u, dt, dirO, hook, lam="Koren", scheme="WH"):
    """Transport part of MPFMomStep! (flow.jl:61,69-70,74,89-92) on the oracle with a `hook(u, rhou, f, stage)` standing in for
    the forcing + projection blocks (flow.jl:75-82, 95-106): it may change u in place.  Mirrors api.mom_advect_step(project=...).
    f and u are advanced in place; returns ρu after the corrector."""
    T = st["dtype"]
    a = alloc_cmom(st)
    u0 = u.copy(order="F"); f0 = f.copy(order="F")                                                     # :61
    O.u2rhou(a["rhou"], u0, f0, st["lam_rho"]); O.BC(a["rhou"], st["uBC"], False, st["perdir"])       # :69
    O.advectVOFrhouu(f0, a["ff"], a["alpha"], a["nhat"], u0, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u,
                     a["alpha"], a["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO)  # :70
    f0[...] = (f0 + f) * T(0.5)                                                                        # :74
    hook(u, a["rhou"], f0, "predictor")                                                                # :75-82
    f0[...] = f                                                                                        # :89
    O.u2rhou(a["rhou"], u0, f, st["lam_rho"]); O.BC(a["rhou"], st["uBC"], False, st["perdir"])        # :91
    O.advectVOFrhouu(f, a["ff"], a["alpha"], a["nhat"], u, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u0,
                     a["alpha"], a["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO)  # :92
    hook(u, a["rhou"], f, "corrector")                                                                 # :95-106
    return a["rhou"]


def _oproject(u, pois, dt):
    """myproject! on the oracle: both inproject! methods (Poisson / MultiLevelPoisson, flow.jl:343-353)."""
    return O.ml_myproject(u, pois, dt) if isinstance(pois, O.MultiLevelPoisson) else O.myproject(u, pois, dt)


def oracle_mom_step_forcing(st, a, f, u, dt, dirO, mu, lam_mu, eta, g, lam="Koren", scheme="WH", pois=None):
    """MPFMomStep! with its explicit forcing (flow.jl:60-107) on the oracle; pois=None leaves the Poisson solve out (:81-82,:105-106),
    pois = O.Poisson(p, a["mu0"], a["Phi"]) runs update!(b); myproject!(a,b[,1/2]); BC! at its two places.
    `a` is the persistent array set of alloc_cmom (the reference's aliasing: uStar≡n̂, dilaU≡α, r≡flow.f, Φ≡flow.σ, fbuffer≡fᶠ) plus
    "mu0"; f and u are advanced in place.  Mirrors api.mom_step_forcing."""
    T = st["dtype"]
    lr, uBC, pd = st["lam_rho"], st["uBC"], st["perdir"]
    u0 = u.copy(order="F"); f0 = f.copy(order="F")                                                      # :61
    O.u2rhou(a["rhou"], u0, f0, lr); O.BC(a["rhou"], uBC, False, pd)                                    # :69
    O.advectVOFrhouu(f0, a["ff"], a["alpha"], a["nhat"], u0, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u,
                     a["alpha"], a["drho"], lr, lam, scheme, uBC, pd, False, dirO)                      # :70
    a["mu0"][...] = 1                                                                                   # :73
    f0[...] = (f0 + f) * T(0.5)                                                                         # :74
    O.viscSurfTenrhou(a["r"], u, a["Phi"], f0, a["alpha"], a["nhat"], a["ff"], lam_mu, mu, lr, eta, pd)  # :75
    O.u2rhou(a["nhat"], u0, f, lr)                                                                      # :76
    O.updateU(u, a["rhou"], a["nhat"], a["r"], dt, f0, lr, g, 0.5)                                      # :77
    O.BC(u, uBC, False, pd)                                                                             # :79
    O.updateL(a["mu0"], f0, lr, pd)                                                                     # :80
    if pois is not None:
        pois.update(); _oproject(u, pois, T(0.5) * T(dt)); O.BC(u, uBC, False, pd)                    # :81-82
    f0[...] = f                                                                                         # :89
    O.u2rhou(a["rhou"], u0, f, lr); O.BC(a["rhou"], uBC, False, pd)                                     # :91
    O.advectVOFrhouu(f, a["ff"], a["alpha"], a["nhat"], u, u, dt, a["cbar"], a["rhou"], a["r"], a["Phi"], a["rhouf"], a["nhat"], u0,
                     a["alpha"], a["drho"], lr, lam, scheme, uBC, pd, False, dirO)                      # :92
    a["mu0"][...] = 1                                                                                   # :96
    O.viscSurfTenrhou(a["r"], u, a["Phi"], f, a["alpha"], a["nhat"], a["ff"], lam_mu, mu, lr, eta, pd)   # :98
    O.u2rhou(a["nhat"], u0, f, lr)                                                                      # :99
    u0[...] = u                                                                                         # :100
    O.updateU(u, a["rhou"], a["nhat"], a["r"], dt, f, lr, g, 1.0)                                       # :101
    O.BC(u, uBC, False, pd)                                                                             # :103
    O.updateL(a["mu0"], f, lr, pd)                                                                      # :104
    if pois is not None:
        pois.update(); _oproject(u, pois, T(1) * T(dt)); O.BC(u, uBC, False, pd)                      # :105-106
    return a["rhou"]
