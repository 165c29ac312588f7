"""Host-side mirror of the reference's interface for the VOF + CMOM advection path.

Reference entry points mirrored here (file:line in /root/reference):
  cVOF                    src/cVOF.jl:18-89          TwoPhaseSimulation   src/InterfaceAdvection.jl:64-85
  advect! / advectVOF!    src/advection.jl:17-78     sim_step!            src/InterfaceAdvection.jl:100-103
  advectfq!/advectVOFρuu! src/flow.jl:157-210        MPFMomStep!          src/flow.jl:60-109 (transport part)
  u2ρu!/ρu2u!             src/VOFutil.jl:198-211     MPCFL                src/flow.jl:262-281
  applyVOF!, BCf!         src/VOFutil.jl:9-37,64-75  BC!                  WaterLily

Arrays are torch CUDA tensors with the reference's Julia shapes -- scalar (N1+2, N2+2[, N3+2]), vector (..., D) --
and COLUMN-MAJOR strides (first index fastest, component slowest), i.e. byte-identical to the CuArrays the Julia
glue hands to the same C ABI.  Indices into them are 0-based here; directions / dirO / perdir stay 1-based.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import LIMITERS, NORMAL_SCHEMES, Context, IfadvError, Report

_DT = {torch.float32: "float32", torch.float64: "float64"}


# ---- column-major tensors ------------------------------------------------------------------------------------
def jl_empty(shape: Sequence[int], dtype=torch.float32, device="cuda") -> torch.Tensor:
    t = torch.empty(tuple(reversed(shape)), dtype=dtype, device=device)
    return t.permute(*reversed(range(len(shape))))


def jl_zeros(shape, dtype=torch.float32, device="cuda") -> torch.Tensor:
    t = jl_empty(shape, dtype, device)
    t.zero_()
    return t


def from_numpy(a: np.ndarray, device="cuda") -> torch.Tensor:
    """Fortran-ordered numpy array -> column-major device tensor (same bytes)."""
    a = np.asfortranarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)
    return t.permute(*reversed(range(a.ndim)))


def to_numpy(t: torch.Tensor) -> np.ndarray:
    """Column-major device tensor -> Fortran-ordered numpy array."""
    _check(t)
    return np.asfortranarray(t.permute(*reversed(range(t.dim()))).contiguous().cpu().numpy().T)


def _check(t: torch.Tensor):
    if not t.is_cuda:
        raise IfadvError("the B200 path needs CUDA tensors; there is no CPU fallback")
    exp, s = [], 1
    for n in t.shape:
        exp.append(s)
        s *= n
    if tuple(t.stride()) != tuple(exp):
        raise IfadvError("tensor is not column-major (use jl_zeros / from_numpy)")


def _copy(dst: torch.Tensor, src: torch.Tensor):
    """copyto!(dst, src) for two column-major fields: one flat vectorised copy of the underlying storage
    (torch's strided elementwise kernel is ~2x slower on permuted views)."""
    _check(dst)
    _check(src)
    dst.as_strided((dst.numel(),), (1,)).copy_(src.as_strided((src.numel(),), (1,)))


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    _check(t)
    return t.data_ptr()


_contexts = {}


def context_for(f: torch.Tensor) -> Context:
    if not f.is_cuda:
        raise IfadvError("the B200 path needs CUDA tensors; there is no CPU fallback")
    key = (tuple(f.shape), f.dtype, f.device.index)
    if key not in _contexts:
        _contexts[key] = Context(tuple(f.shape), _DT[f.dtype], f.device.index or 0)
    return _contexts[key]


def register_context(shape, dtype, device_index, ctx: Context):
    """Bind a pre-built context (e.g. a z-slab context, Context(..., slab=...)) to every field of this local shape / dtype / device."""
    _contexts[(tuple(shape), dtype, device_index)] = ctx


def _stream(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise IfadvError("the B200 path needs CUDA tensors; there is no CPU fallback")
    return torch.cuda.current_stream(t.device).cuda_stream


def _ns(s):
    return NORMAL_SCHEMES[s] if isinstance(s, str) else int(s)


def _lim(s):
    return LIMITERS[s] if isinstance(s, str) else int(s)


# ---- field utilities -------------------------------------------------------------------------------------------
def BCf(f, perdir=()):
    """BCf!(f;perdir)  (VOFutil.jl:64-75)"""
    context_for(f).bcf(_stream(f), _p(f), perdir)


def BC(a, A, saveexit=False, perdir=()):
    """WaterLily.BC!(a,A,saveexit,perdir) for a constant tuple A"""
    if callable(A):
        raise IfadvError("function-valued uBC is not supported by the B200 path")
    context_for(a[..., 0]).bc_vec(_stream(a), _p(a), A, saveexit, perdir)


def u2rhou(rhou, u, f, lam_rho):
    """u2ρu!(ρu,u,f,λρ)  (VOFutil.jl:208-211)"""
    context_for(f).u2rhou(_stream(f), _p(rhou), _p(u), _p(f), lam_rho)


def rhou2u(u, rhou, f, lam_rho):
    """ρu2u!(u,ρu,f,λρ)  (VOFutil.jl:198-201)"""
    context_for(f).rhou2u(_stream(f), _p(u), _p(rhou), _p(f), lam_rho)


def sum_inside(f) -> float:
    """sum(f[inside(f)]) in Float64 (the mass check of test/maintests.jl:209,215)"""
    return context_for(f).sum_inside(_stream(f), _p(f))


def _cell_centres(shape, dtype, device, origin=None):
    D = len(shape)
    o = origin or (0,) * D
    grids = torch.meshgrid(*[torch.arange(n, device=device, dtype=dtype) for n in shape], indexing="ij")
    # 0-based idx <-> Julia I = idx+1; loc(0,I) = I - 1.5 (+ the global offset of a slab)
    return torch.stack([g - 0.5 + o[k] for k, g in enumerate(grids)], dim=-1)


def applyVOF(f, alpha, nhat, InterfaceSDF: Optional[Callable], origin=None):
    """applyVOF!(f,α,n̂,InterfaceSDF)  (VOFutil.jl:8-37).  InterfaceSDF maps a (..., D) tensor of positions to
    signed distances (dark fluid negative) and is evaluated on the device in f's dtype."""
    if InterfaceSDF is None:
        return
    D = f.dim()
    xc = _cell_centres(f.shape, f.dtype, f.device, origin)
    dx = torch.tensor(0.01, dtype=f.dtype, device=f.device)

    def colmajor(t):
        out = jl_empty(t.shape, f.dtype, f.device)
        out.copy_(t)
        return out

    sc = colmajor(InterfaceSDF(xc).to(f.dtype))
    sp = jl_empty(tuple(f.shape) + (D,), f.dtype, f.device)
    sm = jl_empty(tuple(f.shape) + (D,), f.dtype, f.device)
    for i in range(D):
        e = torch.zeros(D, dtype=f.dtype, device=f.device)
        e[i] = dx
        sp[..., i] = InterfaceSDF(xc + e).to(f.dtype)
        sm[..., i] = InterfaceSDF(xc - e).to(f.dtype)
    context_for(f).apply_vof_samples(_stream(f), _p(f), _p(alpha), _p(nhat), _p(sc), _p(sp), _p(sm))


# ---- structs -----------------------------------------------------------------------------------------------------
class cVOF:
    """cVOF(N; T, InterfaceSDF, μ, λμ, λρ, η, normalScheme, perdir)  (src/cVOF.jl:45-89)"""

    def __init__(self, N, T=torch.float32, InterfaceSDF=None, mu=1e-3, lam_mu=1e-2, lam_rho=1e-3, eta=None, normalScheme="WH",
                 perdir=(), device="cuda", origin=None):
        D = len(N)
        Ng = tuple(n + 2 for n in N)
        Nv = Ng + (D,)
        self.D, self.N, self.T = D, tuple(N), T
        self.f = jl_zeros(Ng, T, device)
        self.f.fill_(1)
        self.alpha = jl_zeros(Ng, T, device)
        self.nhat = jl_zeros(Nv, T, device)
        self.cbar = jl_zeros(Ng, torch.int8, device)
        self.perdir = tuple(perdir)
        if InterfaceSDF is not None:
            applyVOF(self.f, self.alpha, self.nhat, InterfaceSDF, origin)
            BCf(self.f, self.perdir)
        self.f0 = self.f.clone(memory_format=torch.preserve_format)
        self.ff = jl_zeros(Ng, T, device)
        self.rhou = jl_zeros(Nv, T, device)
        self.rhouf = jl_zeros(Nv, T, device)
        self.drho = jl_zeros(Nv, T, device)
        self.drho.fill_(1)
        self.eta = None if (eta is None or eta == 0) else eta
        self.mu = None if (mu is None or mu == 0) else mu
        self.lam_rho, self.lam_mu = lam_rho, lam_mu
        self.normalScheme = normalScheme


class Flow:
    """The subset of WaterLily.Flow the path touches: u, u⁰, f (momentum scratch r), σ, Δt, ν, g, uBC, exitBC,
    perdir, λ  (SURVEY App. A)."""

    def __init__(self, N, uBC, T=torch.float32, u0fn=None, dt=0.25, nu=0.0, g=None, exitBC=False, perdir=(), lam="Koren",
                 device="cuda"):
        D = len(N)
        Ng = tuple(n + 2 for n in N)
        self.D, self.N, self.T = D, tuple(N), T
        self.u = jl_zeros(Ng + (D,), T, device)
        if u0fn is not None:  # apply!(u0, u): component i sampled at its face centre loc(i,I)
            xc = _cell_centres(Ng, T, device)
            for i in range(D):
                e = torch.zeros(D, dtype=T, device=device)
                e[i] = 0.5
                self.u[..., i] = u0fn(i + 1, xc - e).to(T)
        self.uBC, self.exitBC, self.perdir, self.lam = tuple(uBC), bool(exitBC), tuple(perdir), lam
        BC(self.u, self.uBC, self.exitBC, self.perdir)
        self.u0 = self.u.clone(memory_format=torch.preserve_format)
        self.f = jl_zeros(Ng + (D,), T, device)
        self.sigma = jl_zeros(Ng, T, device)
        self.p = jl_zeros(Ng, T, device)
        self.mu0 = jl_zeros(Ng + (D,), T, device)  # μ₀: the projection's face coefficients (updateL!, flow.jl:254-259)
        self.mu0.fill_(1)
        self.dt = [float(dt)]  # Δt
        self.nu, self.g = nu, g


class TwoPhaseSimulation:
    """TwoPhaseSimulation(dims, u_BC, L; T, λμ, λρ, η, InterfaceSDF, λ, normalScheme, kwargs...)
    (src/InterfaceAdvection.jl:64-85).  psolver: "Poisson" (the Jacobi-PCG psolver! of flow.jl:300), "MultiLevelPoisson" (WaterLily's
    default, inproject!'s second method) or None (no projection object); bodies stay with WaterLily (out of scope)."""

    def __init__(self, dims, u_BC, L, T=torch.float32, lam_mu=1e-2, lam_rho=1e-3, eta=None, InterfaceSDF=None, lam="Koren",
                 normalScheme="WH", U=None, dt=0.25, nu=0.0, g=None, u0=None, perdir=(), exitBC=False, device="cuda", psolver=None):
        self.L = L
        self.U = U if U is not None else math.sqrt(sum(float(x) ** 2 for x in u_BC))
        self.flow = Flow(dims, u_BC, T=T, u0fn=u0, dt=dt, nu=nu, g=g, exitBC=exitBC, perdir=perdir, lam=lam, device=device)
        self.intf = cVOF(dims, T=T, InterfaceSDF=InterfaceSDF, mu=nu, lam_mu=lam_mu, lam_rho=lam_rho, eta=eta,
                         normalScheme=normalScheme, perdir=perdir, device=device)
        self.flow.dt[-1] = min(self.flow.dt[-1], MPCFL(self.flow, self.intf))  # InterfaceAdvection.jl:81
        # psolver=Poisson: WaterLily's Simulation(...; psolver=Poisson) -- the solver the reference's psolver! is written for
        # (flow.jl:300); psolver=MultiLevelPoisson: WaterLily's default, inproject!'s second method (flow.jl:343-347); None: no projection
        # object (velocities prescribed, or a `project` hook of the caller's)
        if psolver in ("Poisson", Poisson):
            self.pois = Poisson(self.flow.p, self.flow.mu0, self.flow.sigma, perdir=perdir)
        elif psolver in ("MultiLevelPoisson", MultiLevelPoisson):
            self.pois = MultiLevelPoisson(self.flow.p, self.flow.mu0, self.flow.sigma, perdir=perdir)
        else:
            self.pois = None
        self.body = None


# ---- the hot path ------------------------------------------------------------------------------------------------------
def _dirO(flow, D):
    n = len(flow.dt)
    return tuple((n + i) % D + 1 for i in range(1, D + 1))  # advection.jl:22, flow.jl:163


def _report(status, rep: Report):
    if status > 0:  # the reference prints and carries on (advection.jl:150-166); the exploding-divergence case has already been
        #             raised as IfadvError by the library (status -5, advection.jl:160,180)
        which = "max" if status & 1 else "min"
        val = rep.maxf - 1 if status & 1 else -rep.minf
        idx = tuple(rep.argmax) if status & 1 else tuple(rep.argmin)
        print(f"|∇⋅u⁰| = {rep.div_u0:+13.8f}, |∇⋅u| = {rep.div_u:+13.8f}")
        print(f"ERROR: {which} VOF @ {idx} ∉ [0,1] @ direction {rep.dir}, Δf = {val}")
    return status


def advectVOF(f, ff, alpha, nhat, u, u0, dt, cbar, rhouf, lam_rho, normalScheme="WH", perdir=(), dirO=None, check=True,
              want_rhouf=True):
    """advectVOF!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρuf,λρ,normalScheme; perdir,dirO)  (src/advection.jl:34-78)"""
    D = f.dim()
    if dirO is None:  # the reference shuffles when no order is given (advection.jl:34)
        dirO = tuple(int(x) + 1 for x in np.random.permutation(D))
    rep = Report() if check else None
    st = context_for(f).advect_vof(_stream(f), _p(f), _p(ff), _p(alpha), _p(nhat), _p(u), _p(u0), dt, _p(cbar), _p(rhouf), lam_rho,
                                   _ns(normalScheme), perdir, dirO, 0 if want_rhouf else _lib.IFADV_NO_RHOUF, rep)
    return _report(st, rep) if check else st


def advect(a: Flow, c: cVOF, f=None, u1=None, u2=None, dt=None, check=True, want_rhouf=True):
    """advect!(a,c,f,u¹,u²,dt)  (src/advection.jl:17-23).  want_rhouf=False skips the un-normalised mass flux ρuf that advect! leaves
    behind for its callers (IFADV_NO_RHOUF: D·s bytes per cell and step less, plus the fill!(ρuf,0) pass)."""
    f = c.f if f is None else f
    u1 = a.u0 if u1 is None else u1
    u2 = a.u if u2 is None else u2
    dt = a.dt[-1] if dt is None else dt
    return advectVOF(f, c.ff, c.alpha, c.nhat, u1, u2, dt, c.cbar, c.rhouf, c.lam_rho, c.normalScheme, perdir=a.perdir,
                     dirO=_dirO(a, a.D), check=check, want_rhouf=want_rhouf)


def advectVOFrhouu(f, ff, alpha, nhat, u, u0, dt, cbar, rhou, r, Phi, rhouf, uStar, uOld, dilaU, drho, lam_rho, lam="Koren",
                   normalScheme="WH", uBC=(0, 0, 0), perdir=(), exitBC=False, dirO=None, check=True):
    """advectVOFρuu!(f,fᶠ,α,n̂,u,u⁰,Δt,c̄,ρu,r,Φ,ρuf,uStar,uOld,dilaU,dρ,λρ,λ,normalScheme,uBC; perdir,exitBC,dirO)
    (src/flow.jl:165-210)"""
    D = f.dim()
    if callable(uBC):
        raise IfadvError("function-valued uBC is not supported by the B200 path")
    if dirO is None:
        dirO = tuple(int(x) + 1 for x in np.random.permutation(D))
    rep = Report() if check else None
    st = context_for(f).advect_vof_rhouu(_stream(f), _p(f), _p(ff), _p(alpha), _p(nhat), _p(u), _p(u0), dt, _p(cbar), _p(rhou), _p(r),
                                         _p(Phi), _p(rhouf), _p(uStar), _p(uOld), _p(dilaU), _p(drho), lam_rho, _lim(lam),
                                         _ns(normalScheme), uBC, perdir, exitBC, dirO, rep)
    return _report(st, rep) if check else st


def advectfq(a: Flow, c: cVOF, f=None, u1=None, u2=None, u0=None, dt=None, check=True):
    """advectfq!(a,c,f,u¹,u²,u⁰,dt)  (src/flow.jl:157-164): binds r≡flow.f, Φ≡flow.σ, uStar≡n̂, dilaU≡α."""
    f = c.f if f is None else f
    u1 = a.u0 if u1 is None else u1
    u2 = a.u if u2 is None else u2
    u0 = a.u if u0 is None else u0
    dt = a.dt[-1] if dt is None else dt
    return advectVOFrhouu(f, c.ff, c.alpha, c.nhat, u1, u2, dt, c.cbar, c.rhou, a.f, a.sigma, c.rhouf, c.nhat, u0, c.alpha, c.drho,
                          c.lam_rho, a.lam, c.normalScheme, a.uBC, perdir=a.perdir, exitBC=a.exitBC, dirO=_dirO(a, a.D), check=check)


def MPCFL(a: Flow, c: cVOF, dt_max=1.0, safety=0.8) -> float:
    """MPCFL(a,c)  (src/flow.jl:262-281)"""
    gnorm = 0.0
    if a.g is not None:
        gnorm = math.sqrt(sum(x * x for x in _gvec(a.g, a.D, sum(a.dt))))
    return context_for(c.f).mpcfl(_stream(c.f), _p(a.u), nu=a.nu, mu=c.mu or 0.0, lam_mu=c.lam_mu, lam_rho=c.lam_rho, eta=c.eta or 0.0,
                                  gnorm=gnorm, dt_max=dt_max, safety=safety)


def u2rhou_advectfq(a: Flow, c: cVOF, f_src, f, u1, u2, uOld, dt=None, check=False):
    """Fused  f .= f_src; u2ρu!(c.ρu,uOld,f,λρ); BC!(c.ρu,…); advectfq!(a,c,f,u¹,u²,uOld,dt)  -- the three calls MPFMomStep! makes
    back to back at src/flow.jl:61,69-70 and :89-92 (ifadv_u2rhou_advect_vof_rhouu; bit-identical to the separate calls)."""
    dt = a.dt[-1] if dt is None else dt
    rep = Report() if check else None
    st = context_for(f).u2rhou_advect_vof_rhouu(_stream(f), _p(f_src), _p(f), _p(c.ff), _p(a.sigma), _p(u1), _p(u2), dt, _p(c.cbar),
                                                _p(c.rhou), _p(a.f), _p(c.rhouf), _p(uOld), _p(c.drho), c.lam_rho, _lim(a.lam),
                                                _ns(c.normalScheme), a.uBC, a.perdir, a.exitBC, _dirO(a, a.D), rep)
    return _report(st, rep) if check else st


_side_streams = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    key = dev.index or 0
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def mom_advect_step(a: Flow, c: cVOF, dt=None, project: Optional[Callable] = None, check=False, fused=True, overlap=True):
    """The transport half of MPFMomStep! (src/flow.jl:61,69-70,74,89-92): two u2ρu!+BC!+advectfq! groups and the
    midpoint f⁰.  This is one "advection step" of the benchmark metric; MPCFL is separate.
    fused=True issues each group through the fused entry point (same results bit for bit, fewer passes).
    overlap=True (only without a `project` hook, i.e. with prescribed velocities): the three bandwidth-bound field operations
    of the step run on a second stream underneath the issue-bound sweeps -- u⁰←u (:61) beside the predictor, which reads u for
    both velocity arguments (u⁰≡u at that point), and the midpoint (:74) + f⁰←f (:89) beside the first two sweeps of the
    corrector, which does not write f before its last sweep (ifadv_defer_f_writes_until).  Same operations, same results."""
    dt = a.dt[-1] if dt is None else dt
    ctx, s = context_for(c.f), _stream(c.f)
    if overlap and fused and project is None:
        main = torch.cuda.current_stream(c.f.device)
        side = _side_stream(c.f.device)
        side.wait_stream(main)                                        # u and f are final on the main stream
        with torch.cuda.stream(side):
            _copy(a.u0, a.u)                                          # :61
            ev_u0 = side.record_event()
        u2rhou_advectfq(a, c, c.f, c.f0, a.u, a.u, a.u, dt, check=check)    # :61 (f⁰←f), :69, :70 with u⁰≡u
        ev_g1 = main.record_event()
        with torch.cuda.stream(side):
            side.wait_event(ev_g1)
            ctx.axpby(side.cuda_stream, _p(c.f0), 0.5, _p(c.f0), 0.5, _p(c.f))  # :74
            _copy(c.f0, c.f)                                          # :89
            ev_f0 = side.record_event()
        main.wait_event(ev_u0)                                        # the corrector reads u⁰ as uOld
        ctx.defer_f_writes_until(ev_f0.cuda_event)
        u2rhou_advectfq(a, c, c.f, c.f, a.u, a.u, a.u0, dt, check=check)    # :91, :92
        return
    _copy(a.u0, a.u)                                                  # :61
    if fused:
        u2rhou_advectfq(a, c, c.f, c.f0, a.u0, a.u, a.u, dt, check=check)   # :61 (f⁰←f), :69, :70
    else:
        _copy(c.f0, c.f)                                              # :61
        u2rhou(c.rhou, a.u0, c.f0, c.lam_rho)
        BC(c.rhou, a.uBC, a.exitBC, a.perdir)                         # :69
        advectfq(a, c, c.f0, a.u0, a.u, a.u, dt, check=check)         # :70
    ctx.axpby(s, _p(c.f0), 0.5, _p(c.f0), 0.5, _p(c.f))               # :74
    if project is not None:
        project(a, c, "predictor")                                    # :75-82 (WaterLily side)
        ctx.exchange_planes(s, _p(a.u), a.D, a.u.element_size())      # z-slab: the changed u needs its ghost planes (no-op on one GPU)
    _copy(c.f0, c.f)                                                  # :89
    if fused:
        u2rhou_advectfq(a, c, c.f, c.f, a.u, a.u, a.u0, dt, check=check)    # :91, :92
    else:
        u2rhou(c.rhou, a.u0, c.f, c.lam_rho)
        BC(c.rhou, a.uBC, a.exitBC, a.perdir)                         # :91
        advectfq(a, c, c.f, a.u, a.u, a.u0, dt, check=check)          # :92
    if project is not None:
        project(a, c, "corrector")                                    # :95-106 (WaterLily side)
        ctx.exchange_planes(s, _p(a.u), a.D, a.u.element_size())


# ---- explicit forcing between advection and projection ---------------------------------------------------------
def viscSurfTenrhou(r, u, Phi, f, alpha, nhat, fbuffer, lam_mu, mu, lam_rho, eta, perdir=()):
    """viscSurfTenρu!(r,u,Φ,f,α,n̂,fbuffer,λμ,μ,λρ,η;perdir)  (src/flow.jl:113-117; visc! :120-152; surfTen! src/surfaceTension.jl:8-21).
    mu / eta = None stand for `nothing`."""
    return context_for(f).visc_surften_rhou(_stream(f), _p(r), _p(u), _p(Phi), _p(f), _p(alpha), _p(nhat), _p(fbuffer), lam_mu, mu,
                                            lam_rho, eta, perdir)


def _gvec(g, D, t):
    """WaterLily's g(i,x,t) evaluated as a constant vector (the only form accelerate! is restated for)."""
    if g is None:
        return None
    if callable(g):
        return tuple(float(g(i + 1, [0.0] * D, t)) for i in range(D))
    return tuple(float(x) for x in g)


def updateU(u, rhou, rhou0, forcing, dt, f, lam_rho, tNow=0.0, g=None, uBC=None, w=1.0):
    """updateU!(u,ρu,ρu⁰,forcing,dt,f,λρ,tNow,g,uBC,w)  (src/flow.jl:244-252); g: None, a constant vector or g(i,x,t) sampled at x=0."""
    return context_for(f).update_u(_stream(f), _p(u), _p(rhou), _p(rhou0), _p(forcing), dt, _p(f), lam_rho, _gvec(g, f.dim(), tNow), w)


def updateL(mu0, f, lam_rho, perdir=(), fill_one=False):
    """updateL!(μ₀,f,λρ;perdir)  (src/flow.jl:254-259); fill_one folds the preceding fill!(μ₀,1) (flow.jl:73,96) into the pass."""
    return context_for(f).update_l(_stream(f), _p(mu0), _p(f), lam_rho, perdir, fill_one)


def _ustar_top_planes(nhat, uBC, perdir, exitBC):
    """In the reference n̂ ≡ u★ leaves advectfq! with BC! applied (flow.jl:197): plane N_d of component d holds uBC[d].  visc! reads
    exactly that plane as `fFace` (f2face!+BCv! never write it).  The B200 sweeps do not materialise u★, so the mirror writes the
    D planes the reference's side effect leaves behind (exitBC: plane N of component 1 is skipped by BC!, like there)."""
    D = nhat.dim() - 1
    for d in range(D):
        if (d + 1) in perdir or (exitBC and d == 0):
            continue
        idx = [slice(None)] * D + [d]
        idx[d] = -1
        nhat[tuple(idx)] = float(uBC[d])


def mom_step_forcing(a: Flow, c: cVOF, dt=None, project: Optional[Callable] = None, check=False):
    """MPFMomStep! with its explicit forcing (src/flow.jl:60-107): transport (B200 sweeps), viscSurfTenρu!, updateU!, BC!, updateL! for
    the predictor (weight 1/2) and the corrector.  `project(a, c, stage)` stands for update!(b); myproject!(a,b[,1/2]) (WaterLily's
    Poisson solve, out of scope) and is called at its place (:82, :106); without it u stays unprojected."""
    dt = a.dt[-1] if dt is None else dt
    ctx, s = context_for(c.f), _stream(c.f)
    t1 = sum(a.dt); t0 = t1 - dt; tm = t1 - dt / 2
    _copy(a.u0, a.u)                                                          # :61
    u2rhou_advectfq(a, c, c.f, c.f0, a.u0, a.u, a.u, dt, check=check)          # :61 (f⁰←f), :69, :70
    ctx.axpby(s, _p(c.f0), 0.5, _p(c.f0), 0.5, _p(c.f))                        # :74
    _ustar_top_planes(c.nhat, a.uBC, a.perdir, a.exitBC)
    viscSurfTenrhou(a.f, a.u, a.sigma, c.f0, c.alpha, c.nhat, c.ff, c.lam_mu, c.mu, c.lam_rho, c.eta, a.perdir)  # :75
    u2rhou(c.nhat, a.u0, c.f, c.lam_rho)                                       # :76
    updateU(a.u, c.rhou, c.nhat, a.f, dt, c.f0, c.lam_rho, tm, a.g, a.uBC, 0.5)  # :77
    BC(a.u, a.uBC, a.exitBC, a.perdir)                                         # :79
    updateL(a.mu0, c.f0, c.lam_rho, a.perdir, fill_one=True)                   # :73, :80
    if project is not None:
        project(a, c, "predictor")                                             # :81-82
        BC(a.u, a.uBC, a.exitBC, a.perdir)
    _copy(c.f0, c.f)                                                           # :89
    u2rhou_advectfq(a, c, c.f, c.f, a.u, a.u, a.u0, dt, check=check)           # :91, :92
    _ustar_top_planes(c.nhat, a.uBC, a.perdir, a.exitBC)
    viscSurfTenrhou(a.f, a.u, a.sigma, c.f, c.alpha, c.nhat, c.ff, c.lam_mu, c.mu, c.lam_rho, c.eta, a.perdir)   # :98
    u2rhou(c.nhat, a.u0, c.f, c.lam_rho)                                       # :99
    _copy(a.u0, a.u)                                                           # :100
    updateU(a.u, c.rhou, c.nhat, a.f, dt, c.f, c.lam_rho, t1, a.g, a.uBC)      # :101
    BC(a.u, a.uBC, a.exitBC, a.perdir)                                         # :103
    updateL(a.mu0, c.f, c.lam_rho, a.perdir, fill_one=True)                    # :96, :104
    if project is not None:
        project(a, c, "corrector")                                             # :105-106
        BC(a.u, a.uBC, a.exitBC, a.perdir)


# ---- pressure projection ----------------------------------------------------------------------------------------
class Poisson:
    """WaterLily.Poisson(x,L,z;perdir): x ≡ Flow.p, L ≡ Flow.μ₀, z ≡ Flow.σ are the caller's arrays; D, iD, ϵ, r its own; n collects
    the iteration counts like the reference's `p.n`."""

    def __init__(self, x, L, z, perdir=()):
        self.x, self.L, self.z, self.perdir = x, L, z, tuple(perdir)
        Ng, T, dev = tuple(x.shape), x.dtype, x.device
        self.D, self.iD, self.eps, self.r = (jl_zeros(Ng, T, dev) for _ in range(4))
        self.n, self.r2 = [], []
        update(self)


def update(b):
    """update!(b::Poisson) = set_diag!(D,iD,L); update!(b::MultiLevelPoisson) = set_diag! + restrictL! down the levels (flow.jl:81,105)."""
    if isinstance(b, MultiLevelPoisson):
        return b._ctx.ml_update(b._h, _stream(b.x))
    return context_for(b.x).poisson_update(_stream(b.x), _p(b.D), _p(b.iD), _p(b.L))


def psolver(b: Poisson, tol=None, itmx=6000):
    """psolver!(p;tol=50eps(T),itmx=6e3) (src/flow.jl:300-326) -> iterations."""
    n, r2 = context_for(b.x).psolver(_stream(b.x), _p(b.x), _p(b.eps), _p(b.r), _p(b.z), _p(b.L), _p(b.D), _p(b.iD), b.perdir, tol, itmx)
    b.n.append(n); b.r2.append(r2)
    return n


def myproject(a: Flow, b, w=1.0):
    """myproject!(a,b,w) (src/flow.jl:328-347): dt = T(w)·last(a.Δt); z ← ∇·u, x ← x·dt, psolver! (b::Poisson) or
    solver!(b;tol=1e-4,itmx=200) (b::MultiLevelPoisson), u -= L ∂x, x ← x/dt."""
    T = a.u.dtype
    dt = float(torch.tensor(w, dtype=T) * torch.tensor(a.dt[-1], dtype=T))
    if isinstance(b, MultiLevelPoisson):
        n, r2 = b._ctx.ml_myproject(b._h, _stream(b.x), _p(a.u), dt)
        b.n.append(n); b.r2.append(r2)
        return n
    n, r2 = context_for(b.x).myproject(_stream(b.x), _p(a.u), _p(b.x), _p(b.eps), _p(b.r), _p(b.z), _p(b.L), _p(b.D), _p(b.iD), dt, b.perdir)
    b.n.append(n); b.r2.append(r2)
    return n


def project_with(b) -> Callable:
    """The `project(a, c, stage)` hook of MPFMomStep / mom_step_forcing for a Poisson / MultiLevelPoisson: update!(b); myproject!(a,b[,1/2])
    (flow.jl:81-82,105-106)."""
    def hook(a, c, stage):
        update(b)
        myproject(a, b, 0.5 if stage == "predictor" else 1.0)
    return hook


class _DevView:
    """A device array the library owns, exposed through __cuda_array_interface__ (no copy; valid while its owner lives)."""

    def __init__(self, ptr, shape_c, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape_c), "typestr": "<f4" if dtype == torch.float32 else "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class MultiLevelPoisson:
    """WaterLily.MultiLevelPoisson(x,L,z;maxlevels=10,perdir): geometric multigrid on the caller's x ≡ Flow.p, L ≡ Flow.μ₀, z ≡ Flow.σ
    (level 1); the coarser levels and every level's D, iD, ϵ, r live in the library's handle (ifadv_ml_*).  The solver behind
    inproject!(a,b::MultiLevelPoisson,dt) (src/flow.jl:343-347) and WaterLily's default `psolver`.  n collects the cycle counts like the
    reference's `ml.n`.  level(l, name) is a column-major view of a level's array (name: L, D, iD, x, eps, r, z)."""
    _NAMES = {"L": 0, "D": 1, "iD": 2, "x": 3, "eps": 4, "r": 5, "z": 6}

    def __init__(self, x, L, z, perdir=(), maxlevels=10):
        self.x, self.L, self.z, self.perdir = x, L, z, tuple(perdir)
        self._ctx = context_for(x)
        self._h = self._ctx.ml_create(_stream(x), _p(x), _p(L), _p(z), self.perdir, maxlevels)
        if self._ctx.ml_levels(self._h) <= 2:  # WaterLily's constructor asserts length(levels) > 2
            self._ctx.ml_destroy(self._h)
            self._h = None
            raise IfadvError("MultiLevelPoisson requires size=a2ⁿ, where n>2")
        self.n, self.r2 = [], []

    def __del__(self):
        try:
            if self._h is not None:
                self._ctx.ml_destroy(self._h)
        except Exception:
            pass

    @property
    def levels(self) -> int:
        return self._ctx.ml_levels(self._h)

    def level(self, l, name) -> torch.Tensor:
        ptr, ng = self._ctx.ml_level_array(self._h, l, self._NAMES[name])
        D = self.x.dim()
        shape = tuple(ng[:D]) + ((D,) if name == "L" else ())
        t = torch.as_tensor(_DevView(ptr, tuple(reversed(shape)), self.x.dtype), device=self.x.device)
        return t.permute(*reversed(range(len(shape))))


def Vcycle(ml: MultiLevelPoisson):
    """Vcycle!(ml) (WaterLily MultiLevelPoisson.jl)."""
    return ml._ctx.ml_vcycle(ml._h, _stream(ml.x))


def smooth(ml: MultiLevelPoisson, level=0):
    """smooth!(ml.levels[level+1]) = pcg!(p;it=6)."""
    return ml._ctx.ml_smooth(ml._h, _stream(ml.x), level)


def residual(ml: MultiLevelPoisson):
    """residual!(ml): r = z - A x on level 1, mean removed."""
    return ml._ctx.ml_residual(ml._h, _stream(ml.x))


def solver(ml: MultiLevelPoisson, tol=1e-4, itmx=32):
    """solver!(ml;tol=1e-4,itmx=32) -> V-cycles."""
    n, r2 = ml._ctx.ml_solver(ml._h, _stream(ml.x), tol, itmx)
    ml.n.append(n); ml.r2.append(r2)
    return n


# ---- post-processing --------------------------------------------------------------------------------------------
class LevelSet:
    """LevelSet(sim) (src/redistaning.jl:8-29): ϕ = 2f-1 in sim.intf.f⁰'s storage, ϕ⁰ ≡ α, ϕini ≡ fᶠ, L ≡ flow.σ -- no extra memory."""

    def __init__(self, sim: TwoPhaseSimulation):
        c, a = sim.intf, sim.flow
        self.phi, self.phi0, self.phi_ini, self.L, self.perdir = c.f0, c.alpha, c.ff, a.sigma, tuple(c.perdir)
        context_for(c.f).levelset_init(_stream(c.f), _p(self.phi), _p(self.phi_ini), _p(c.f))


def computeL(L, phi, phi_ini, perdir=()):
    """computeL!(L,ϕ,ϕini;perdir) (src/redistaning.jl:67-87)."""
    return context_for(phi).redist_compute_l(_stream(phi), _p(L), _p(phi), _p(phi_ini), perdir)


def redistaningStage(phi, phi0, phi_ini, L, dtau, alpha, perdir=()):
    """_redistaningStage!(ϕ,ϕ⁰,ϕini,L,dτ,α;perdir) (src/redistaning.jl:31-34)."""
    return context_for(phi).redist_stage(_stream(phi), _p(phi), _p(phi0), _p(phi_ini), _p(L), dtau, alpha, perdir)


def redistaning(ls: LevelSet, d=5, dtau=0.5, perdir=()):
    """redistaning!(ls; d, dτ, perdir) (src/redistaning.jl:44-57)."""
    return context_for(ls.phi).redistance(_stream(ls.phi), _p(ls.phi), _p(ls.phi0), _p(ls.phi_ini), _p(ls.L), d, dtau, perdir)


def metrics(u, f, lam_rho, U=None, g=None, statWL=None):
    """Σ over inside(f) of ρkeI, ρgh and ρuI(i) (src/metrics.jl:15-17,25,49-51): (kinetic energy, potential energy, momentum[D])."""
    return context_for(f).metrics(_stream(f), _p(u), _p(f), lam_rho, U, g, statWL)


def enstrophy(omega, f):
    """Σ EnsI(I,ω) over the inside cells (src/metrics.jl:34-41); f only selects the context (grid, dtype)."""
    return context_for(f).enstrophy(_stream(f), _p(omega))


def MPFMomStep(a: Flow, b, c: cVOF, d=None, dt=None, project: Optional[Callable] = None, check=False, forcing=False):
    """Transport part of MPFMomStep!(a,b,c,d)  (src/flow.jl:60-109).

    Lines 61, 69-70, 74, 89-92 and 108 run on the B200 kernels.  The forcing (viscSurfTenρu!, updateU!) and the
    pressure projection (myproject!) stay on WaterLily's backend; `project(a, c, stage)` is the hook
    where a caller plugs them in.  Without it velocities are prescribed: u is left untouched between the stages.
    forcing=True runs the reference's whole sequence except the Poisson solve -- viscSurfTenρu!, updateU!, BC!, updateL! on the B200
    kernels too (mom_step_forcing) -- and `project` then stands for update!(b); myproject! only."""
    dt = a.dt[-1] if dt is None else dt
    if project is None and forcing and isinstance(b, (Poisson, MultiLevelPoisson)):
        project = project_with(b)  # the reference's own projection on the B200 kernels
    if forcing:
        mom_step_forcing(a, c, dt, project=project, check=check)
    else:
        mom_advect_step(a, c, dt, project=project, check=check)
    a.dt.append(min(MPCFL(a, c), 1.2 * dt))                           # :108


def sim_step(sim: TwoPhaseSimulation, t_end: Optional[float] = None, project=None, check=False, forcing=False):
    """sim_step!(sim[,t_end])  (src/InterfaceAdvection.jl:100-103 + WaterLily's generic loop)"""
    if t_end is None:
        return MPFMomStep(sim.flow, sim.pois, sim.intf, sim.body, project=project, check=check, forcing=forcing)
    while sim_time(sim) < t_end:
        MPFMomStep(sim.flow, sim.pois, sim.intf, sim.body, project=project, check=check, forcing=forcing)


def sim_time(sim: TwoPhaseSimulation) -> float:
    return sum(sim.flow.dt[:-1]) * sim.U / sim.L
