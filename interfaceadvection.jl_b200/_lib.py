"""ctypes binding of libifadv_b200.so (include/ifadv.h).  There is NO fallback: if the CUDA library is
missing or no CUDA device is present every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IFADV_LIB") or os.path.join(_HERE, "libifadv_b200.so")  # IFADV_LIB: A/B builds of the same ABI (csrc/Makefile OUT=)

NORMAL_SCHEMES = {"WH": 0, "WY": 1, "Column": 2, "PCD": 3, "SLIC": 4, "MYC": 5, "Y": 6, "CD": 7, "XYLIC": 8}
LIMITERS = {"upwind": 0, "minmod": 1, "Koren": 2, "vanAlbada1": 3, "Sweby": 4, "superbee": 5, "TVDcen": 6, "TVDdown": 7,
            "quick": 8, "vanLeer": 9, "cds": 10}
IFADV_NO_RHOUF = 1


class IfadvError(RuntimeError):
    pass


class Report(C.Structure):
    _fields_ = [("maxf", C.c_double), ("minf", C.c_double), ("argmax", C.c_int64 * 3), ("argmin", C.c_int64 * 3),
                ("dir", C.c_int), ("status", C.c_int), ("div_u0", C.c_double), ("div_u", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IfadvError(f"{LIB_PATH} is missing: build it with `make -C interfaceadvection.jl_b200/csrc -j8` "
                             "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, dbl, i32, u32, i64p = C.c_void_p, C.c_double, C.c_int, C.c_uint, C.POINTER(C.c_int64)
        i32p, dblp, rep = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(Report)
        L.ifadv_version.restype = C.c_char_p
        L.ifadv_last_error.restype = C.c_char_p
        L.ifadv_last_error.argtypes = [vp]
        L.ifadv_launch_count.restype = C.c_int64
        L.ifadv_launch_count.argtypes = [vp]
        L.ifadv_profile.argtypes = [vp, i32]
        L.ifadv_profile_read.argtypes = [vp, dblp, i64p]
        L.ifadv_profile_read_dirs.argtypes = [vp, dblp, i64p]
        L.ifadv_create.argtypes = [C.POINTER(vp), i32, i64p, i32, i32]
        L.ifadv_destroy.argtypes = [vp]
        L.ifadv_advect_vof.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, dbl, vp, vp, dbl, i32, u32, i32p, i32, rep]
        L.ifadv_advect_vof_rhouu.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, i32, i32,
                                             dblp, u32, i32, i32p, rep]
        L.ifadv_u2rhou_advect_vof_rhouu.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, dbl, i32, i32, dblp, u32,
                                                    i32, i32p, rep]
        L.ifadv_u2rhou.argtypes = [vp, vp, vp, vp, vp, dbl]
        L.ifadv_rhou2u.argtypes = [vp, vp, vp, vp, vp, dbl]
        L.ifadv_bc_vec.argtypes = [vp, vp, vp, dblp, i32, u32]
        L.ifadv_bcf.argtypes = [vp, vp, vp, u32]
        L.ifadv_axpby.argtypes = [vp, vp, vp, dbl, vp, dbl, vp]
        L.ifadv_mpcfl.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl, dbl, dbl, dblp]
        L.ifadv_sum_inside.argtypes = [vp, vp, vp, dblp]
        L.ifadv_apply_vof_samples.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.ifadv_mom_advect_step_host.argtypes = [vp, vp, vp, vp, dbl, dbl, i32, i32, dblp, u32, i32p, rep]
        L.ifadv_host_step_bytes.argtypes = [vp, i64p, i64p, i32p]
        L.ifadv_defer_f_writes_until.argtypes = [vp, vp]
        L.ifadv_visc_surften_rhou.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, u32]
        L.ifadv_update_u.argtypes = [vp, vp, vp, vp, vp, vp, dbl, vp, dbl, dblp, dbl]
        L.ifadv_update_l.argtypes = [vp, vp, vp, vp, dbl, u32, i32]
        L.ifadv_levelset_init.argtypes = [vp, vp, vp, vp, vp]
        L.ifadv_redist_compute_l.argtypes = [vp, vp, vp, vp, vp, u32]
        L.ifadv_redist_stage.argtypes = [vp, vp, vp, vp, vp, vp, dbl, dbl, u32]
        L.ifadv_redistance.argtypes = [vp, vp, vp, vp, vp, vp, dbl, dbl, u32]
        L.ifadv_metrics.argtypes = [vp, vp, vp, vp, dbl, dblp, dblp, dblp, dblp]
        L.ifadv_enstrophy.argtypes = [vp, vp, vp, dblp]
        L.ifadv_check_nan.argtypes = [vp, vp]
        L.ifadv_poisson_update.argtypes = [vp, vp, vp, vp, vp]
        L.ifadv_psolver.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, dbl, i32, i32p, dblp]
        L.ifadv_myproject.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, u32, i32p, dblp]
        L.ifadv_ml_create.argtypes = [vp, C.POINTER(vp), vp, vp, vp, vp, u32, i32]
        L.ifadv_ml_destroy.argtypes = [vp]
        L.ifadv_ml_levels.argtypes = [vp]
        L.ifadv_ml_level_array.argtypes = [vp, i32, i32, C.POINTER(vp), i64p]
        L.ifadv_ml_update.argtypes = [vp, vp]
        L.ifadv_ml_residual.argtypes = [vp, vp]
        L.ifadv_ml_vcycle.argtypes = [vp, vp]
        L.ifadv_ml_smooth.argtypes = [vp, vp, i32]
        L.ifadv_ml_solver.argtypes = [vp, vp, dbl, i32, i32p, dblp]
        L.ifadv_ml_myproject.argtypes = [vp, vp, vp, dbl, i32p, dblp]
        L.ifadv_create_slab.argtypes = [C.POINTER(vp), i64p, i32, i32, vp, i32, i32, i32, i32]
        L.ifadv_slab_info.argtypes = [vp, i32p, i32p, i32p, i32p, i64p]
        L.ifadv_exchange_planes.argtypes = [vp, vp, vp, i32, i32]
        L.ifadv_slab_p2p.argtypes = [vp]
        L.ifadv_nccl_unique_id.argtypes = [C.c_char_p]
        L.ifadv_nccl_comm_init.argtypes = [C.POINTER(vp), i32, C.c_char_p, i32, i32]
        L.ifadv_nccl_comm_destroy.argtypes = [vp]
        _lib = L
    return _lib


def perdir_mask(perdir) -> int:
    m = 0
    for j in perdir:
        m |= 1 << (int(j) - 1)
    return m


def _d3(v, D):
    v = [float(x) for x in v][:D]
    return (C.c_double * 3)(*(v + [0.0] * (3 - D)))


def _i3(v, D):
    v = [int(x) for x in v][:D]
    return (C.c_int * 3)(*(v + [0] * (3 - D)))


class Context:
    """ifadv_ctx: one per (device, grid, dtype).  Ng = array extents including ghosts (N .+ 2)."""

    def __init__(self, Ng, dtype: str, device: int = 0, slab=None):
        """slab: None, or dict(comm=<ncclComm_t as int>, rank, nranks, G, per_z) -> ifadv_create_slab (Ng = LOCAL extents)."""
        self.D = len(Ng)
        self.Ng = tuple(int(n) for n in Ng)
        self.dtype = {"float32": 0, "float64": 1}[dtype]
        self._h = C.c_void_p()
        ng = (C.c_int64 * 3)(*(list(self.Ng) + [1] * (3 - self.D)))
        if slab is None:
            rc = lib().ifadv_create(C.byref(self._h), self.D, ng, self.dtype, int(device))
        else:
            rc = lib().ifadv_create_slab(C.byref(self._h), ng, self.dtype, int(device), C.c_void_p(slab["comm"]), int(slab["rank"]),
                                         int(slab["nranks"]), int(slab["G"]), int(bool(slab["per_z"])))
        if rc != 0:
            raise IfadvError(f"ifadv_create{'_slab' if slab else ''} failed ({rc}): a CUDA device is required, there is no CPU fallback")

    def slab_info(self):
        """-> dict(kz0, kz1, lower, upper, bytes_sent): owned planes [kz0, kz1) (1-based), neighbour ranks, exchange volume"""
        a, b, lo, up, n = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int64(0)
        self._chk(lib().ifadv_slab_info(self._h, C.byref(a), C.byref(b), C.byref(lo), C.byref(up), C.byref(n)))
        return dict(kz0=a.value, kz1=b.value, lower=lo.value, upper=up.value, bytes_sent=int(n.value),
                    transport="cuda-ipc peer-to-peer (copy engines)" if lib().ifadv_slab_p2p(self._h) else "nccl send/recv")

    def exchange_planes(self, stream, field, ncomp, elem_bytes):
        return self._chk(lib().ifadv_exchange_planes(self._h, stream, field, int(ncomp), int(elem_bytes)))

    def close(self):
        if self._h:
            lib().ifadv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            msg = lib().ifadv_last_error(self._h).decode()
            if rc == -1:
                raise IfadvError("NaN!")  # error("NaN!"), src/advection.jl:148
            if rc == -5:
                raise IfadvError(msg)  # error("divergence, …, is exploding!"), src/advection.jl:160,180
            raise IfadvError(f"ifadv error {rc}: {msg}")
        return rc

    @property
    def launches(self) -> int:
        return int(lib().ifadv_launch_count(self._h))

    def profile(self, enable: bool):
        return self._chk(lib().ifadv_profile(self._h, int(bool(enable))))

    def profile_read(self):
        """-> ((ms_standard, n_standard), (ms_fused_first, n_fused_first))"""
        ms, n = (C.c_double * 2)(), (C.c_int64 * 2)()
        self._chk(lib().ifadv_profile_read(self._h, ms, n))
        return (ms[0], int(n[0])), (ms[1], int(n[1]))

    def profile_read_dirs(self):
        """-> {"x": ms/launch, "x_fused": ..} accumulated by the profile_read calls since the last call"""
        ms, n = (C.c_double * 6)(), (C.c_int64 * 6)()
        self._chk(lib().ifadv_profile_read_dirs(self._h, ms, n))
        out = {}
        for j, nm in enumerate("xyz"):
            for fu in (0, 1):
                if n[2 * j + fu]:
                    out[nm + ("_fused_first" if fu else "")] = ms[2 * j + fu] / n[2 * j + fu]
        return out

    def advect_vof(self, stream, f, ff, alpha, nhat, u, u0, dt, cbar, rhouf, lam_rho, scheme, perdir, dirO, flags=0, report=None):
        r = C.byref(report) if report is not None else None
        return self._chk(lib().ifadv_advect_vof(self._h, stream, f, ff, alpha, nhat, u, u0, float(dt), cbar, rhouf, float(lam_rho),
                                                int(scheme), perdir_mask(perdir), _i3(dirO, self.D), int(flags), r))

    def advect_vof_rhouu(self, stream, f, ff, alpha, nhat, u, u0, dt, cbar, rhou, r_, Phi, rhouf, uStar, uOld, dilaU, drho, lam_rho,
                         limiter, scheme, uBC, perdir, exitBC, dirO, report=None):
        r = C.byref(report) if report is not None else None
        return self._chk(lib().ifadv_advect_vof_rhouu(self._h, stream, f, ff, alpha, nhat, u, u0, float(dt), cbar, rhou, r_, Phi, rhouf,
                                                      uStar, uOld, dilaU, drho, float(lam_rho), int(limiter), int(scheme),
                                                      _d3(uBC, self.D), perdir_mask(perdir), int(bool(exitBC)), _i3(dirO, self.D), r))

    def u2rhou_advect_vof_rhouu(self, stream, f_src, f, ff, Phi, u, u0, dt, cbar, rhou, r_, rhouf, uOld, drho, lam_rho, limiter, scheme,
                                uBC, perdir, exitBC, dirO, report=None):
        r = C.byref(report) if report is not None else None
        return self._chk(lib().ifadv_u2rhou_advect_vof_rhouu(self._h, stream, f_src, f, ff, Phi, u, u0, float(dt), cbar, rhou, r_, rhouf,
                                                             uOld, drho, float(lam_rho), int(limiter), int(scheme),
                                                             _d3(uBC, self.D), perdir_mask(perdir), int(bool(exitBC)), _i3(dirO, self.D), r))

    def u2rhou(self, stream, rhou, u, f, lam_rho):
        return self._chk(lib().ifadv_u2rhou(self._h, stream, rhou, u, f, float(lam_rho)))

    def rhou2u(self, stream, u, rhou, f, lam_rho):
        return self._chk(lib().ifadv_rhou2u(self._h, stream, u, rhou, f, float(lam_rho)))

    def bc_vec(self, stream, a, A, saveexit, perdir):
        return self._chk(lib().ifadv_bc_vec(self._h, stream, a, _d3(A, self.D), int(bool(saveexit)), perdir_mask(perdir)))

    def bcf(self, stream, f, perdir):
        return self._chk(lib().ifadv_bcf(self._h, stream, f, perdir_mask(perdir)))

    def axpby(self, stream, out, a, x, b, y):
        return self._chk(lib().ifadv_axpby(self._h, stream, out, float(a), x, float(b), y))

    def mpcfl(self, stream, u, nu=0.0, mu=0.0, lam_mu=1e-2, lam_rho=1e-3, eta=0.0, gnorm=0.0, dt_max=1.0, safety=0.8) -> float:
        out = C.c_double()
        self._chk(lib().ifadv_mpcfl(self._h, stream, u, nu, mu, lam_mu, lam_rho, eta, gnorm, dt_max, safety, C.byref(out)))
        return out.value

    def sum_inside(self, stream, f) -> float:
        out = C.c_double()
        self._chk(lib().ifadv_sum_inside(self._h, stream, f, C.byref(out)))
        return out.value

    def apply_vof_samples(self, stream, f, alpha, nhat, sc, sp, sm):
        return self._chk(lib().ifadv_apply_vof_samples(self._h, stream, f, alpha, nhat, sc, sp, sm))

    def visc_surften_rhou(self, stream, r, u, Phi, f, alpha, nhat, fbuffer, lam_mu, mu, lam_rho, eta, perdir):
        """viscSurfTenρu! (ifadv_visc_surften_rhou); mu / eta None or 0 stand for `nothing`."""
        return self._chk(lib().ifadv_visc_surften_rhou(self._h, stream, r, u, Phi, f, alpha, nhat, fbuffer, float(lam_mu), float(mu or 0.0),
                                                       float(lam_rho), float(eta or 0.0), perdir_mask(perdir)))

    def update_u(self, stream, u, rhou, rhou0, forcing, dt, f, lam_rho, g=None, w=1.0):
        gv = None if g is None else _d3(g, self.D)
        return self._chk(lib().ifadv_update_u(self._h, stream, u, rhou, rhou0, forcing, float(dt), f, float(lam_rho), gv, float(w)))

    def update_l(self, stream, mu0, f, lam_rho, perdir, fill_one=False):
        return self._chk(lib().ifadv_update_l(self._h, stream, mu0, f, float(lam_rho), perdir_mask(perdir), int(bool(fill_one))))

    def levelset_init(self, stream, phi, phi_ini, f):
        return self._chk(lib().ifadv_levelset_init(self._h, stream, phi, phi_ini, f))

    def redist_compute_l(self, stream, L, phi, phi_ini, perdir):
        return self._chk(lib().ifadv_redist_compute_l(self._h, stream, L, phi, phi_ini, perdir_mask(perdir)))

    def redist_stage(self, stream, phi, phi0, phi_ini, L, dtau, alpha, perdir):
        return self._chk(lib().ifadv_redist_stage(self._h, stream, phi, phi0, phi_ini, L, float(dtau), float(alpha), perdir_mask(perdir)))

    def redistance(self, stream, phi, phi0, phi_ini, L, d, dtau, perdir):
        return self._chk(lib().ifadv_redistance(self._h, stream, phi, phi0, phi_ini, L, float(d), float(dtau), perdir_mask(perdir)))

    def metrics(self, stream, u, f, lam_rho, U=None, g=None, statWL=None):
        """(Σρke, Σρgh, [Σρu_i]) over inside(f) (ifadv_metrics)."""
        out = (C.c_double * 5)()
        t3 = lambda v: None if v is None else _d3(v, self.D)
        self._chk(lib().ifadv_metrics(self._h, stream, u, f, float(lam_rho), t3(U), t3(g), t3(statWL), out))
        return out[0], out[1], [out[2 + i] for i in range(self.D)]

    def enstrophy(self, stream, omega) -> float:
        out = C.c_double()
        self._chk(lib().ifadv_enstrophy(self._h, stream, omega, C.byref(out)))
        return out.value

    def poisson_update(self, stream, D, iD, L):
        return self._chk(lib().ifadv_poisson_update(self._h, stream, D, iD, L))

    def psolver(self, stream, x, eps, r, z, L, D, iD, perdir, tol=None, itmx=6000):
        """(iterations, last r₂) of psolver!(p;tol,itmx) (ifadv_psolver); tol=None: 50eps(T)."""
        n, r2 = C.c_int(0), C.c_double(0.0)
        self._chk(lib().ifadv_psolver(self._h, stream, x, eps, r, z, L, D, iD, perdir_mask(perdir), -1.0 if tol is None else float(tol),
                                      int(itmx), C.byref(n), C.byref(r2)))
        return n.value, r2.value

    def myproject(self, stream, u, x, eps, r, z, L, D, iD, dt, perdir):
        """(iterations, last r₂) of myproject!(a,b,w), dt = T(w)·last(a.Δt) (ifadv_myproject)."""
        n, r2 = C.c_int(0), C.c_double(0.0)
        self._chk(lib().ifadv_myproject(self._h, stream, u, x, eps, r, z, L, D, iD, float(dt), perdir_mask(perdir), C.byref(n), C.byref(r2)))
        return n.value, r2.value

    # ---- WaterLily.MultiLevelPoisson (ifadv_ml_*): the handle is a plain integer owned by the caller (api.MultiLevelPoisson) ----
    def ml_create(self, stream, x, L, z, perdir, maxlevels=10) -> int:
        h = C.c_void_p()
        self._chk(lib().ifadv_ml_create(self._h, C.byref(h), stream, x, L, z, perdir_mask(perdir), int(maxlevels)))
        return h.value

    def ml_destroy(self, h):
        lib().ifadv_ml_destroy(C.c_void_p(h))

    def ml_levels(self, h) -> int:
        return lib().ifadv_ml_levels(C.c_void_p(h))

    def ml_level_array(self, h, level, which):
        """(device pointer, extents incl. ghosts) of a level's array; which: 0 L, 1 D, 2 iD, 3 x, 4 ϵ, 5 r, 6 z."""
        ptr, ng = C.c_void_p(), (C.c_int64 * 3)()
        self._chk(lib().ifadv_ml_level_array(C.c_void_p(h), int(level), int(which), C.byref(ptr), ng))
        return ptr.value, tuple(ng)

    def ml_update(self, h, stream):
        return self._chk(lib().ifadv_ml_update(C.c_void_p(h), stream))

    def ml_residual(self, h, stream):
        return self._chk(lib().ifadv_ml_residual(C.c_void_p(h), stream))

    def ml_vcycle(self, h, stream):
        return self._chk(lib().ifadv_ml_vcycle(C.c_void_p(h), stream))

    def ml_smooth(self, h, stream, level=0):
        return self._chk(lib().ifadv_ml_smooth(C.c_void_p(h), stream, int(level)))

    def ml_solver(self, h, stream, tol=1e-4, itmx=32):
        """(cycles, last r₂) of solver!(ml;tol,itmx) (ifadv_ml_solver)."""
        n, r2 = C.c_int(0), C.c_double(0.0)
        self._chk(lib().ifadv_ml_solver(C.c_void_p(h), stream, float(tol), int(itmx), C.byref(n), C.byref(r2)))
        return n.value, r2.value

    def ml_myproject(self, h, stream, u, dt):
        """(cycles, last r₂) of myproject!(a,b::MultiLevelPoisson,w), dt = T(w)·last(a.Δt) (ifadv_ml_myproject)."""
        n, r2 = C.c_int(0), C.c_double(0.0)
        self._chk(lib().ifadv_ml_myproject(C.c_void_p(h), stream, u, float(dt), C.byref(n), C.byref(r2)))
        return n.value, r2.value

    def defer_f_writes_until(self, event):
        """One-shot: the next CMOM advect call waits for `event` (cudaEvent_t handle) before its first write to f."""
        return self._chk(lib().ifadv_defer_f_writes_until(self._h, event))

    def check_nan(self, stream):
        """Raises IfadvError("NaN!") if any sweep since the last check / reporting call produced a NaN in f (ifadv_check_nan)."""
        return self._chk(lib().ifadv_check_nan(self._h, stream))

    def host_step_bytes(self):
        """(h2d_bytes, d2h_bytes, slabs) of the last mom_advect_step_host call (ifadv_host_step_bytes)."""
        a, b, n = C.c_int64(0), C.c_int64(0), C.c_int(0)
        lib().ifadv_host_step_bytes(self._h, C.byref(a), C.byref(b), C.byref(n))
        return int(a.value), int(b.value), int(n.value)

    def mom_advect_step_host(self, f_host, u_host, rhou_host, dt, lam_rho, limiter, scheme, uBC, perdir, dirO, report=None):
        r = C.byref(report) if report is not None else None
        return self._chk(lib().ifadv_mom_advect_step_host(self._h, f_host, u_host, rhou_host, float(dt), float(lam_rho), int(limiter),
                                                          int(scheme), _d3(uBC, self.D), perdir_mask(perdir), _i3(dirO, self.D), r))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().ifadv_nccl_unique_id(buf)
    if rc != 0:
        raise IfadvError(f"ifadv_nccl_unique_id failed ({rc})")
    return buf.raw


def nccl_comm_init(nranks: int, uid: bytes, rank: int, device: int) -> int:
    comm = C.c_void_p()
    rc = lib().ifadv_nccl_comm_init(C.byref(comm), int(nranks), C.create_string_buffer(uid, 128), int(rank), int(device))
    if rc != 0:
        raise IfadvError(f"ifadv_nccl_comm_init failed ({rc})")
    return comm.value


def nccl_comm_destroy(comm: int):
    lib().ifadv_nccl_comm_destroy(C.c_void_p(comm))
