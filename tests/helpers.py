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
