"""ctypes driver for the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module.  Arrays are numpy arrays in Fortran (column-major) order with the Julia shapes of the reference:
scalar fields (N1+2, N2+2[, N3+2]); vector fields (..., D) with the component index slowest.
Indices, directions and dirO are 1-based like the Julia reference; perdir is a tuple of 1-based directions.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NORMAL_SCHEMES = {"WH": 0, "WY": 1, "Column": 2, "PCD": 3, "SLIC": 4, "MYC": 5, "Y": 6, "CD": 7, "XYLIC": 8}
LIMITERS = {"upwind": 0, "minmod": 1, "Koren": 2, "vanAlbada1": 3, "Sweby": 4, "superbee": 5, "TVDcen": 6, "TVDdown": 7,
            "quick": 8, "vanLeer": 9, "cds": 10}


class FillReport(C.Structure):
    _fields_ = [("maxf", C.c_double), ("minf", C.c_double), ("argmax", C.c_int64 * 3), ("argmin", C.c_int64 * 3),
                ("dir", C.c_int), ("status", C.c_int), ("div_u0", C.c_double), ("div_u", C.c_double)]


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_omp.so with the committed Makefile (g++ only, a few seconds)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_omp.so"))
    if not need:
        srcs = [os.path.join(_HERE, n) for n in ("oracle.cpp", "oracle_core.hpp", "oracle_fields.hpp", "oracle_flow.hpp", "oracle_forcing.hpp", "oracle_post.hpp", "oracle_poisson.hpp")]
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(os.path.getmtime(os.path.join(_HERE, n)) < newest for n in ("liboracle.so", "liboracle_omp.so"))
    if need:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


_libs = {}


def lib(omp: bool = False):
    name = "liboracle_omp.so" if omp else "liboracle.so"
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _libs[name] = C.CDLL(path)
    return _libs[name]


def mask(perdir) -> int:
    m = 0
    for j in perdir:
        m |= 1 << (int(j) - 1)
    return m


def _dt(a) -> int:
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.float64:
        return 1
    raise TypeError(a.dtype)


def _p(a):
    if a is None:
        return None
    assert a.flags.f_contiguous, "oracle arrays must be Fortran-ordered"
    return C.c_void_p(a.ctypes.data)


def _ng(f):
    D = f.ndim
    return D, (C.c_int64 * 3)(*(list(f.shape) + [1] * (3 - D)))


def _i3(I):
    I = list(I)
    return (C.c_int64 * 3)(*(I + [1] * (3 - len(I))))


def zeros(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


def farr(x, dtype=None):
    return np.asfortranarray(np.array(x, dtype=dtype))


# ---- scalar KAT entry points ---------------------------------------------------------------------
def getIntercept(n, g, dtype=np.float64) -> float:
    out = C.c_double()
    nn = (C.c_double * 3)(*(list(map(float, n)) + [0.0] * (3 - len(n))))
    lib().orc_get_intercept(0 if dtype == np.float32 else 1, len(n), nn, C.c_double(float(g)), C.byref(out))
    return out.value


def getVolumeFraction(n, b, dtype=np.float64) -> float:
    out = C.c_double()
    nn = (C.c_double * 3)(*(list(map(float, n)) + [0.0] * (3 - len(n))))
    lib().orc_get_volume_fraction(0 if dtype == np.float32 else 1, len(n), nn, C.c_double(float(b)), C.byref(out))
    return out.value


def limiter(name, u, c, d, dtype=np.float64) -> float:
    out = C.c_double()
    lib().orc_limiter(0 if dtype == np.float32 else 1, LIMITERS[name], C.c_double(u), C.c_double(c), C.c_double(d), C.byref(out))
    return out.value


# ---- field entry points --------------------------------------------------------------------------
def normal(scheme, f, nhat, I):
    D, ng = _ng(f)
    lib().orc_normal(_dt(f), D, ng, NORMAL_SCHEMES[scheme], _p(f), _p(nhat), _i3(I))


def getVOFFlux_face(ff, f, alpha, nhat, dl, d, IFace, rhouf, lr):
    D, ng = _ng(f)
    lib().orc_vof_flux_face(_dt(f), D, ng, _p(ff), _p(f), _p(alpha), _p(nhat), C.c_double(dl), int(d), _i3(IFace), _p(rhouf),
                            C.c_double(lr))


def BCf(f, perdir=()):
    D, ng = _ng(f)
    lib().orc_bcf(_dt(f), D, ng, _p(f), mask(perdir))


def BCv1D(f, d, perdir=()):
    D, ng = _ng(f)
    lib().orc_bcv1d(_dt(f), D, ng, _p(f), int(d), mask(perdir))


def BCv(v, perdir=()):
    D, ng = _ng(v[..., 0])
    lib().orc_bcv(_dt(v), D, ng, _p(v), mask(perdir))


def BCVOF(f, alpha, nhat, perdir=()):
    D, ng = _ng(f)
    lib().orc_bcvof(_dt(f), D, ng, _p(f), _p(alpha), _p(nhat), mask(perdir))


def BC(a, A, saveexit=False, perdir=()):
    """WaterLily.BC!(a,A,saveexit,perdir) for a constant tuple A."""
    D, ng = _ng(a[..., 0])
    AA = (C.c_double * 3)(*(list(map(float, A)) + [0.0] * (3 - len(A))))
    lib().orc_bc_vec(_dt(a), D, ng, _p(a), AA, int(bool(saveexit)), mask(perdir))


def cleanWisp(f):
    D, ng = _ng(f)
    lib().orc_clean_wisp(_dt(f), D, ng, _p(f))


def u2rhou(rhou, u, f, lr, omp=False):
    D, ng = _ng(f)
    lib(omp).orc_u2rhou(_dt(f), D, ng, _p(rhou), _p(u), _p(f), C.c_double(lr))


def rhou2u(u, rhou, f, lr):
    D, ng = _ng(f)
    lib().orc_rhou2u(_dt(f), D, ng, _p(u), _p(rhou), _p(f), C.c_double(lr))


def f2face(fFace, fCen, perdir=()):
    D, ng = _ng(fCen)
    lib().orc_f2face(_dt(fCen), D, ng, _p(fFace), _p(fCen), mask(perdir))


def applyVOF(f, alpha, nhat, sdf):
    """applyVOF!(f,α,n̂,InterfaceSDF) (VOFutil.jl:9-37); `sdf` maps an array of shape (..., D) of
    positions to signed distances and is evaluated here in f's dtype like the Julia closure would be."""
    D, ng = _ng(f)
    T = f.dtype.type
    idx = np.indices(f.shape).astype(f.dtype)  # 0-based index i  <->  Julia I=i+1, centre loc(0,I)=I-1.5
    xc = np.stack([idx[k] + T(1) - T(1.5) for k in range(D)], axis=-1).astype(f.dtype)
    dx = T(0.01)
    sc = np.asfortranarray(np.asarray(sdf(xc), dtype=f.dtype))
    sp = zeros(f.shape + (D,), f.dtype)
    sm = zeros(f.shape + (D,), f.dtype)
    for i in range(D):
        e = np.zeros(D, dtype=f.dtype)
        e[i] = dx
        sp[..., i] = np.asarray(sdf((xc + e).astype(f.dtype)), dtype=f.dtype)
        sm[..., i] = np.asarray(sdf((xc - e).astype(f.dtype)), dtype=f.dtype)
    lib().orc_apply_vof_samples(_dt(f), D, ng, _p(f), _p(alpha), _p(nhat), _p(sc), _p(sp), _p(sm))


def reconstructInterface(f, alpha, nhat, scheme="WH", perdir=()):
    D, ng = _ng(f)
    lib().orc_reconstruct_interface(_dt(f), D, ng, _p(f), _p(alpha), _p(nhat), NORMAL_SCHEMES[scheme], mask(perdir))


def getVOFFlux(ff, f, alpha, nhat, u, u0, dt, d, rhouf, lr):
    D, ng = _ng(f)
    lib().orc_get_vof_flux(_dt(f), D, ng, _p(ff), _p(f), _p(alpha), _p(nhat), _p(u), _p(u0), C.c_double(dt), int(d), _p(rhouf),
                           C.c_double(lr))


def advectVOF(f, ff, alpha, nhat, u, u0, dt, cbar, rhouf, lr, scheme="WH", perdir=(), dirO=None, omp=False):
    """advectVOF!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρuf,λρ,normalScheme; perdir,dirO)  -> (status, FillReport)"""
    D, ng = _ng(f)
    dirO = tuple(dirO) if dirO is not None else tuple(range(1, D + 1))
    rep = FillReport()
    st = lib(omp).orc_advect_vof(_dt(f), D, ng, _p(f), _p(ff), _p(alpha), _p(nhat), _p(u), _p(u0), C.c_double(dt),
                                 C.c_void_p(cbar.ctypes.data), _p(rhouf), C.c_double(lr), NORMAL_SCHEMES[scheme], mask(perdir),
                                 (C.c_int * 3)(*(list(dirO) + [0] * (3 - D))), C.byref(rep))
    return st, rep


def advectVOFrhouu(f, ff, alpha, nhat, u, u0, dt, cbar, rhou, r, Phi, rhouf, uStar, uOld, dilaU, drho, lr, limiter="Koren",
                   scheme="WH", uBC=(0, 0, 0), perdir=(), exitBC=False, dirO=None, omp=False):
    """advectVOFρuu!(...) (flow.jl:165)  -> (status, FillReport).  uStar may be nhat, dilaU may be alpha."""
    D, ng = _ng(f)
    dirO = tuple(dirO) if dirO is not None else tuple(range(1, D + 1))
    rep = FillReport()
    A = (C.c_double * 3)(*(list(map(float, uBC))[:D] + [0.0] * (3 - D)))
    st = lib(omp).orc_advect_vof_rhouu(_dt(f), D, ng, _p(f), _p(ff), _p(alpha), _p(nhat), _p(u), _p(u0), C.c_double(dt),
                                       C.c_void_p(cbar.ctypes.data), _p(rhou), _p(r), _p(Phi), _p(rhouf), _p(uStar), _p(uOld),
                                       _p(dilaU), _p(drho), C.c_double(lr), LIMITERS[limiter], NORMAL_SCHEMES[scheme], A,
                                       mask(perdir), int(bool(exitBC)), (C.c_int * 3)(*(list(dirO) + [0] * (3 - D))), C.byref(rep))
    return st, rep


def advectVOFrhouu_sweep(iop, f, ff, alpha, nhat, u, u0, dt, cbar, rhou, r, Phi, rhouf, uStar, uOld, dilaU, drho, lr, limiter="Koren",
                         scheme="WH", uBC=(0, 0, 0), perdir=(), exitBC=False, dirO=None):
    """Directional sweep number `iop` (0-based) of advectVOFρuu! on its own (c̄ is computed with sweep 0)."""
    D, ng = _ng(f)
    rep = FillReport()
    A = (C.c_double * 3)(*(list(map(float, uBC))[:D] + [0.0] * (3 - D)))
    st = lib().orc_advect_vof_rhouu_sweep(_dt(f), D, ng, _p(f), _p(ff), _p(alpha), _p(nhat), _p(u), _p(u0), C.c_double(dt),
                                          C.c_void_p(cbar.ctypes.data), _p(rhou), _p(r), _p(Phi), _p(rhouf), _p(uStar), _p(uOld),
                                          _p(dilaU), _p(drho), C.c_double(lr), LIMITERS[limiter], NORMAL_SCHEMES[scheme], A,
                                          mask(perdir), int(bool(exitBC)), (C.c_int * 3)(*(list(dirO) + [0] * (3 - D))), C.byref(rep),
                                          int(iop))
    return st, rep


def MPCFL(u, sigma, nu=0.0, mu=0.0, lam_mu=1e-2, lam_rho=1e-3, eta=0.0, gnorm=0.0, dt_max=1.0, safety=0.8) -> float:
    D, ng = _ng(sigma)
    out = C.c_double()
    lib().orc_mpcfl(_dt(u), D, ng, _p(u), _p(sigma), C.c_double(nu), C.c_double(mu), C.c_double(lam_mu), C.c_double(lam_rho),
                    C.c_double(eta), C.c_double(gnorm), C.c_double(dt_max), C.c_double(safety), C.byref(out))
    return out.value


def sum_inside(f) -> float:
    D, ng = _ng(f)
    out = C.c_double()
    lib().orc_sum_inside(_dt(f), D, ng, _p(f), C.byref(out))
    return out.value


# ---- explicit forcing -------------------------------------------------------------------------------------------
def getmu(i, j, I, fFace, lam_mu, mu, lam_rho) -> float:
    """getμ(i,j,I,fFace,λμ,μ,λρ) (VOFutil.jl:186-191); i, j, I 1-based."""
    D = fFace.ndim - 1
    ng = (C.c_int64 * 3)(*(list(fFace.shape[:D]) + [1] * (3 - D)))
    out = C.c_double()
    lib().orc_getmu(_dt(fFace), D, ng, _p(fFace), int(i), int(j), _i3(I), C.c_double(lam_mu), C.c_double(mu), C.c_double(lam_rho),
                    C.byref(out))
    return out.value


def getPopinetHeight(I, f, i) -> float:
    """getPopinetHeight(I,f,i) (surfaceTension.jl:67-99); i: signed 1-based direction."""
    D, ng = _ng(f)
    out = C.c_double()
    lib().orc_popinet_height(_dt(f), D, ng, _p(f), _i3(I), int(i), C.byref(out))
    return out.value


def getCurvature(I, f, i) -> float:
    """getCurvature(I,f,i) (surfaceTension.jl:31-65)."""
    D, ng = _ng(f)
    out = C.c_double()
    lib().orc_curvature(_dt(f), D, ng, _p(f), _i3(I), int(i), C.byref(out))
    return out.value


def viscSurfTenrhou(r, u, Phi, f, alpha, nhat, fbuffer, lam_mu, mu, lam_rho, eta, perdir=(), omp=False):
    """viscSurfTenρu!(r,u,Φ,f,α,n̂,fbuffer,λμ,μ,λρ,η;perdir) (flow.jl:113-117); mu / eta = None stand for `nothing`."""
    D, ng = _ng(f)
    lib(omp).orc_visc_surften_rhou(_dt(f), D, ng, _p(r), _p(u), _p(Phi), _p(f), _p(alpha), _p(nhat), _p(fbuffer), C.c_double(lam_mu),
                                   C.c_double(0.0 if mu is None else mu), int(mu is not None), C.c_double(lam_rho),
                                   C.c_double(0.0 if eta is None else eta), int(eta is not None), mask(perdir))


def updateU(u, rhou, rhou0, forcing, dt, f, lam_rho, g=None, w=1.0, omp=False):
    """updateU!(u,ρu,ρu⁰,forcing,dt,f,λρ,tNow,g,uBC,w) (flow.jl:244-252) with a constant gravity vector g (None = nothing)."""
    D, ng = _ng(f)
    gv = None if g is None else (C.c_double * 3)(*(list(map(float, g))[:D] + [0.0] * (3 - D)))
    lib(omp).orc_update_u(_dt(f), D, ng, _p(u), _p(rhou), _p(rhou0), _p(forcing), C.c_double(dt), _p(f), C.c_double(lam_rho), gv,
                          C.c_double(w))


def updateL(mu0, f, lam_rho, perdir=(), omp=False):
    """updateL!(μ₀,f,λρ;perdir) (flow.jl:254-259)."""
    D, ng = _ng(f)
    lib(omp).orc_update_l(_dt(f), D, ng, _p(mu0), _p(f), C.c_double(lam_rho), mask(perdir))


# ---- post-processing: level-set redistancing and metrics ------------------------------------------------------------
def gradphi2(a, b, c, d, e, s, dtype=np.float64) -> float:
    """𝛁ϕᵢ²(a,b,c,d,e,s) (redistaning.jl:119-143)."""
    out = C.c_double()
    lib().orc_gradphi2(0 if np.dtype(dtype) == np.float32 else 1, *(C.c_double(float(v)) for v in (a, b, c, d, e, s)), C.byref(out))
    return out.value


def computeL(L, phi, phi_ini, perdir=()):
    """computeL!(L,ϕ,ϕini;perdir) (redistaning.jl:67-87)."""
    D, ng = _ng(phi)
    lib().orc_compute_l(_dt(phi), D, ng, _p(L), _p(phi), _p(phi_ini), mask(perdir))


def redistaningStage(phi, phi0, phi_ini, L, dtau, alpha, perdir=()):
    """_redistaningStage!(ϕ,ϕ⁰,ϕini,L,dτ,α;perdir) (redistaning.jl:31-34)."""
    D, ng = _ng(phi)
    lib().orc_redist_stage(_dt(phi), D, ng, _p(phi), _p(phi0), _p(phi_ini), _p(L), C.c_double(dtau), C.c_double(alpha), mask(perdir))


def redistaning(phi, phi0, phi_ini, L, d=5, dtau=0.5, perdir=(), omp=False):
    """redistaning!(ls; d, dτ, perdir) (redistaning.jl:44-57) on the LevelSet arrays."""
    D, ng = _ng(phi)
    lib(omp).orc_redistance(_dt(phi), D, ng, _p(phi), _p(phi0), _p(phi_ini), _p(L), C.c_double(d), C.c_double(dtau), mask(perdir))


def _t3(v, D):
    return None if v is None else (C.c_double * 3)(*(list(map(float, v))[:D] + [0.0] * (3 - D)))


def metrics_cell(I, u, f, lam_rho, U=None, g=None, statWL=None):
    """(ρkeI, ρgh, [ρuI(i) for i in 1:D]) at cell I (metrics.jl:15,25,49)."""
    D, ng = _ng(f)
    out = (C.c_double * 5)()
    lib().orc_metrics_cell(_dt(f), D, ng, _p(u), _p(f), C.c_double(lam_rho), _t3(U, D), _t3(g, D), _t3(statWL, D), _i3(I), out)
    return out[0], out[1], [out[2 + i] for i in range(D)]


def metrics_sum(u, f, lam_rho, U=None, g=None, statWL=None):
    """Σ over inside(f) of ρkeI, ρgh, ρuI(i)."""
    D, ng = _ng(f)
    out = (C.c_double * 5)()
    lib().orc_metrics_sum(_dt(f), D, ng, _p(u), _p(f), C.c_double(lam_rho), _t3(U, D), _t3(g, D), _t3(statWL, D), out)
    return out[0], out[1], [out[2 + i] for i in range(D)]


def enstrophy(omega, D, I=None):
    """EnsI at cell I (if given) and Σ EnsI over the inside cells (metrics.jl:34-41); ω: (…,3) in 3-D, scalar field in 2-D."""
    ng = (C.c_int64 * 3)(*(list(omega.shape[:D]) + [1] * (3 - D)))
    cell, tot = C.c_double(), C.c_double()
    lib().orc_enstrophy(_dt(omega), D, ng, _p(omega), _i3(I) if I is not None else None, C.byref(cell), C.byref(tot))
    return (cell.value if I is not None else None), tot.value


# ---- pressure projection: psolver!, inproject!, myproject! on WaterLily's Poisson arrays ------------------------------
class Poisson:
    """WaterLily.Poisson(x,L,z;perdir): x (pressure), L (face coefficients, Flow.μ₀), z (source, Flow.σ) are the caller's arrays;
    D, iD, ϵ, r are its own; update() = set_diag!."""

    def __init__(self, x, L, z, perdir=()):
        self.x, self.L, self.z, self.perdir = x, L, z, tuple(perdir)
        self.D, self.iD, self.eps, self.r = (zeros(x.shape, x.dtype) for _ in range(4))
        self.n = []
        self.update()

    def update(self):
        D, ng = _ng(self.x)
        lib().orc_pois_update(_dt(self.x), D, ng, _p(self.D), _p(self.iD), _p(self.L))


def pois_mult(p: Poisson, x):
    """mult!(p,x): perBC!(x); p.z = A·x on inside(x)."""
    D, ng = _ng(x)
    lib().orc_pois_mult(_dt(x), D, ng, _p(p.z), _p(x), _p(p.L), _p(p.D), mask(p.perdir))
    return p.z


def pois_residual(p: Poisson):
    """residual!(p): r = z - A·x with the mean removed."""
    D, ng = _ng(p.x)
    lib().orc_pois_residual(_dt(p.x), D, ng, _p(p.x), _p(p.r), _p(p.z), _p(p.L), _p(p.D), _p(p.iD), mask(p.perdir))


def psolver(p: Poisson, tol=None, itmx=6000, omp=False):
    """psolver!(p;tol=50eps(T),itmx) (flow.jl:300-326) -> (iterations, last r₂)."""
    D, ng = _ng(p.x)
    r2 = C.c_double()
    n = lib(omp).orc_psolver(_dt(p.x), D, ng, _p(p.x), _p(p.eps), _p(p.r), _p(p.z), _p(p.L), _p(p.D), _p(p.iD), mask(p.perdir),
                             C.c_double(-1.0 if tol is None else tol), int(itmx), C.byref(r2))
    p.n.append(n)
    return n, r2.value


def myproject(u, p: Poisson, dt, omp=False):
    """myproject!(a,b,w) with dt = T(w)·last(a.Δt) (flow.jl:328-347) -> (iterations, last r₂)."""
    D, ng = _ng(p.x)
    r2 = C.c_double()
    n = lib(omp).orc_myproject(_dt(p.x), D, ng, _p(u), _p(p.x), _p(p.eps), _p(p.r), _p(p.z), _p(p.L), _p(p.D), _p(p.iD),
                               mask(p.perdir), C.c_double(float(p.x.dtype.type(dt))), C.byref(r2))
    p.n.append(n)
    return n, r2.value


class MultiLevelPoisson:
    """WaterLily.MultiLevelPoisson(x,L,z;maxlevels=10,perdir): geometric multigrid on the caller's x, L, z (level 1); the coarser
    levels and every level's D, iD, ϵ, r live in the oracle's handle.  level(l, name) returns a numpy view of a level's array."""
    _NAMES = {"L": 0, "D": 1, "iD": 2, "x": 3, "eps": 4, "r": 5, "z": 6}

    def __init__(self, x, L, z, perdir=(), maxlevels=10):
        self.x, self.L, self.z, self.perdir = x, L, z, tuple(perdir)
        D, ng = _ng(x)
        self.D_ = D
        lib().orc_ml_create.restype = C.c_void_p
        self._h = C.c_void_p(lib().orc_ml_create(_dt(x), D, ng, _p(x), _p(L), _p(z), mask(perdir), int(maxlevels)))
        self.n = []

    def __del__(self):
        try:
            lib().orc_ml_destroy(self._h)
        except Exception:
            pass

    @property
    def levels(self) -> int:
        return lib().orc_ml_levels(self._h)

    def level(self, l, name):
        ptr, ng = C.c_void_p(), (C.c_int64 * 3)()
        assert lib().orc_ml_level_array(self._h, int(l), self._NAMES[name], C.byref(ptr), ng) == 0
        shape = tuple(ng[:self.D_]) + ((self.D_,) if name == "L" else ())
        n = int(np.prod(shape))
        ct = C.c_float if self.x.dtype == np.float32 else C.c_double
        buf = (ct * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=self.x.dtype).reshape(shape, order="F")

    def update(self):
        lib().orc_ml_update(self._h)

    def vcycle(self):
        lib().orc_ml_vcycle(self._h)

    def smooth(self, level=0):
        lib().orc_ml_smooth(self._h, int(level))

    def residual(self):
        lib().orc_ml_residual(self._h)

    def solver(self, tol=1e-4, itmx=32):
        """solver!(ml;tol,itmx) -> (V-cycles, last r₂)."""
        r2 = C.c_double()
        n = lib().orc_ml_solver(self._h, C.c_double(tol), int(itmx), C.byref(r2))
        self.n.append(n)
        return n, r2.value


def ml_myproject(u, ml: MultiLevelPoisson, dt):
    """myproject!(a,b::MultiLevelPoisson,w), dt = T(w)·last(a.Δt): solver!(b;tol=1e-4,itmx=200) inside -> (V-cycles, last r₂)."""
    r2 = C.c_double()
    n = lib().orc_ml_myproject(ml._h, _p(u), C.c_double(float(ml.x.dtype.type(dt))), C.byref(r2))
    ml.n.append(n)
    return n, r2.value


def num_threads(omp=True) -> int:
    return lib(omp).orc_num_threads()


def set_num_threads(n: int, omp=True) -> int:
    """omp_set_num_threads on the timing build (an inherited OMP_NUM_THREADS=1, as torchrun exports, would otherwise pin it)."""
    return lib(omp).orc_set_num_threads(int(n))
