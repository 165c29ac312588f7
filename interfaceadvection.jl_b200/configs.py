"""Synthetic input generators for the configurations named in BASELINE.json.

Everything is analytic (no RNG) and written against a tiny array-module shim so that the same formulas run with
numpy (tests, oracle side) and with torch on the device (bench.py at full size).  Index convention: 0-based array
index `idx` <-> Julia I = idx+1; cell centre x = idx - 0.5; lower d-face centre x - 0.5 e_d.
Returned velocity arrays have the Julia shape (N1+2, N2+2[, N3+2], D).
"""
from __future__ import annotations

import math

import numpy as np


class _NP:
    pi = math.pi

    @staticmethod
    def coords(Ng, dtype, origin=None):
        o = origin or (0,) * len(Ng)
        return [g.astype(dtype) - dtype(0.5) + dtype(o[k]) for k, g in enumerate(np.meshgrid(*[np.arange(n) for n in Ng], indexing="ij"))]

    sin, cos, sqrt, minimum, maximum, abs = np.sin, np.cos, np.sqrt, np.minimum, np.maximum, np.abs

    @staticmethod
    def stack(xs):
        return np.asfortranarray(np.stack(xs, axis=-1))

    @staticmethod
    def zeros_like(x):
        return np.zeros_like(x)


def _torch_shim(device):
    import torch

    class _TH:
        pi = math.pi

        @staticmethod
        def coords(Ng, dtype, origin=None):
            o = origin or (0,) * len(Ng)
            gs = torch.meshgrid(*[torch.arange(n, device=device, dtype=dtype) for n in Ng], indexing="ij")
            return [g - 0.5 + o[k] for k, g in enumerate(gs)]

        sin, cos, sqrt, minimum, maximum, abs = torch.sin, torch.cos, torch.sqrt, torch.minimum, torch.maximum, torch.abs

        @staticmethod
        def stack(xs):
            from .api import jl_empty

            out = jl_empty(tuple(xs[0].shape) + (len(xs),), xs[0].dtype, device)
            for i, x in enumerate(xs):
                out[..., i] = x
            return out

        @staticmethod
        def zeros_like(x):
            return torch.zeros_like(x)

    return _TH


def backend(device=None):
    return _NP if device is None else _torch_shim(device)


# ---- velocity fields (sampled at face centres; all discretely solenoidal to round-off) --------------
def rigid_rotation(N, dtype, omega=None, xp=_NP, Nl=None, origin=None):
    """C1: u = -Ω(y-c), v = Ω(x-c) about the domain centre; Ω = 2π/(16 N) so that |u|Δt <= 0.2 with Δt = 1."""
    Ng = tuple(n + 2 for n in (Nl or N))
    X = xp.coords(Ng, dtype, origin)
    om = (2 * math.pi / (16 * N[0])) if omega is None else omega
    cx, cy = N[0] / 2, N[1] / 2
    u = -om * (X[1] - cy)      # x-face: y is the cell-centre coordinate
    v = om * (X[0] - cx)
    return xp.stack([u, v])


def tgv(N, dtype, U=0.25, xp=_NP, Nl=None, origin=None):
    """Taylor-Green field of test/alloctest.jl:16-21 / test/helper.jl:7-9 scaled to the box: zero normal velocity on
    every wall and periodic with the box length, so it serves walls and periodic configs alike."""
    D = len(N)
    Ng = tuple(n + 2 for n in (Nl or N))
    X = xp.coords(Ng, dtype, origin)

    def sc(k, face):  # scaled coordinate (2x - N)π/N of dimension k, on the lower face if `face`
        x = X[k] - (0.5 if face else 0.0)
        return (2 * x - N[k]) * (math.pi / N[k])

    V = U * math.sin(math.pi / N[0]) / math.sin(math.pi / N[1])  # keeps the discrete divergence zero on non-square boxes
    if D == 2:
        u = -U * xp.sin(sc(0, True)) * xp.cos(sc(1, False))
        v = V * xp.cos(sc(0, False)) * xp.sin(sc(1, True))
        return xp.stack([u, v])
    cz = xp.cos(sc(2, False))
    u = -U * xp.sin(sc(0, True)) * xp.cos(sc(1, False)) * cz
    v = V * xp.cos(sc(0, False)) * xp.sin(sc(1, True)) * cz
    return xp.stack([u, v, xp.zeros_like(u)])


def enright(N, dtype, amp=0.2, xp=_NP, Nl=None, origin=None):
    """C2: LeVeque/Enright deformation field as the DISCRETE CURL of a vector potential sampled on cell edges
    (exactly discretely solenoidal).  A = (0, -sin²(πx̂)sin(2πŷ)sin²(πẑ)/π, sin²(πx̂)sin²(πŷ)sin(2πẑ)/π)·N·amp/2."""
    Ng = tuple(n + 2 for n in (Nl or N))
    X = xp.coords(Ng, dtype, origin)
    n = float(N[0])

    def h(k, lower):  # normalised coordinate of dimension k at the cell centre or the lower face/edge
        return (X[k] - (0.5 if lower else 0.0)) / N[k]

    pi = math.pi
    sc = amp / 2 * n / pi

    def Ay(xl, zl):  # A_y lives on y-edges: lower in x and z, centred in y
        return -sc * xp.sin(pi * h(0, xl)) ** 2 * xp.sin(2 * pi * h(1, False)) * xp.sin(pi * h(2, zl)) ** 2

    def Az(xl, yl):  # A_z lives on z-edges: lower in x and y, centred in z
        return sc * xp.sin(pi * h(0, xl)) ** 2 * xp.sin(pi * h(1, yl)) ** 2 * xp.sin(2 * pi * h(2, False))

    def up(a, k):  # a[I+δk] - a[I] with the last plane repeated (ghost faces are overwritten by BC anyway)
        import builtins

        sl_hi = [builtins.slice(None)] * 3
        sl_lo = [builtins.slice(None)] * 3
        sl_hi[k] = builtins.slice(1, None)
        sl_lo[k] = builtins.slice(0, -1)
        d = xp.zeros_like(a)
        d[tuple(sl_lo)] = a[tuple(sl_hi)] - a[tuple(sl_lo)]
        return d

    ay, az = Ay(True, True), Az(True, True)
    # u = ∂y Az - ∂z Ay on x-faces; v = -∂x Az on y-faces (A_x = 0); w = ∂x Ay on z-faces
    u = up(az, 1) - up(ay, 2)
    v = -up(az, 0)
    w = up(ay, 0)
    return xp.stack([u, v, w])


# ---- signed distance functions (dark fluid negative) ------------------------------------------------------------------
def sdf_sphere(centre, R, inside_dark=True):
    def sdf(x):
        r2 = 0
        for k, c in enumerate(centre):
            r2 = r2 + (x[..., k] - c) ** 2
        d = r2 ** 0.5 - R
        return d if inside_dark else -d
    return sdf


def sdf_zalesak(N):
    """Slotted disk of C1: centre (N/2, 3N/4), R = 0.15 N, slot width 0.05 N, slot top at 0.85 N."""
    cx, cy, R, w, top = N / 2, 0.75 * N, 0.15 * N, 0.05 * N, 0.85 * N

    def sdf(x):
        import numpy as _np

        xs, ys = x[..., 0], x[..., 1]
        mod = _np if isinstance(xs, _np.ndarray) else __import__("torch")
        disk = ((xs - cx) ** 2 + (ys - cy) ** 2) ** 0.5 - R
        # slot: |x-cx| < w/2 and y < top  (negative inside the slot)
        slot = mod.maximum(abs(xs - cx) - w / 2, ys - top)
        return mod.maximum(disk, -slot)
    return sdf


def sdf_dambreak(N):
    """C3: water column x < N1/4 ∧ y < N2/2 (box SDF, negative inside)."""
    ax, ay = N[0] / 4, N[1] / 2

    def sdf(x):
        import numpy as _np

        xs, ys = x[..., 0], x[..., 1]
        mod = _np if isinstance(xs, _np.ndarray) else __import__("torch")
        return mod.maximum(xs - ax, ys - ay)
    return sdf


def sdf_sloshing(N, amp=None):
    """C5: free surface y < N2/2 + a·sin(2πx/N1) (vertical = y so z-slabs are balanced)."""
    a = N[1] / 16 if amp is None else amp

    def sdf(x):
        import numpy as _np

        xs, ys = x[..., 0], x[..., 1]
        mod = _np if isinstance(xs, _np.ndarray) else __import__("torch")
        return ys - (N[1] / 2 + a * mod.sin(xs * (2 * math.pi / N[0])))
    return sdf


CONFIGS = {
    # name: (N, dtype, perdir, SDF factory, velocity factory, λρ, CMOM?)
    "C1_zalesak_128": dict(N=(128, 128), dtype="float64", perdir=(), cmom=False, lam_rho=1e-3),
    "C2_enright_256": dict(N=(256, 256, 256), dtype="float32", perdir=(), cmom=False, lam_rho=1e-3),
    "C3_dambreak_512x256x256": dict(N=(512, 256, 256), dtype="float32", perdir=(), cmom=True, lam_rho=1e-3),
    "C4_bubble_512": dict(N=(512, 512, 512), dtype="float32", perdir=(1, 2), cmom=True, lam_rho=1e-3),
    "C5_sloshing_slab_2048x1024x128": dict(N=(2048, 1024, 128), dtype="float32", perdir=(), cmom=True, lam_rho=1e-3),
}


def make_case(name_or_N, dtype=None, device=None, kind=None, Nl=None, origin=None, vel=None):
    """Return dict(N, sdf, u, perdir, lam_rho, cmom) for a named config, or for an explicit (N, kind) at reduced size.
    Nl / origin: sample the velocity on a sub-box of Nl cells whose first interior cell has global index origin+1
    (slab decomposition); N always describes the GLOBAL box the analytic fields are scaled to.
    vel: None = the config's own generator; "enright" = the three-component discrete-curl LeVeque field (every sweep direction
    carries a non-zero flux; compatible with walls and with periodic boxes); "tgv" = Taylor-Green (w ≡ 0 in 3-D)."""
    if isinstance(name_or_N, str):
        cfg = dict(CONFIGS[name_or_N])
        kind = name_or_N.split("_")[0]
        N = cfg["N"]
        dtype = cfg["dtype"] if dtype is None else dtype
    else:
        N = tuple(name_or_N)
        base = {"C1": "C1_zalesak_128", "C2": "C2_enright_256", "C3": "C3_dambreak_512x256x256", "C4": "C4_bubble_512",
                "C5": "C5_sloshing_slab_2048x1024x128"}[kind]
        cfg = dict(CONFIGS[base])
        cfg["N"] = N
        dtype = cfg["dtype"] if dtype is None else dtype
    xp = backend(device)
    if device is None:
        T = np.dtype(dtype).type
    else:
        import torch

        T = getattr(torch, dtype)
    kw = dict(xp=xp, Nl=Nl, origin=origin)
    if kind == "C1":
        sdf, u = sdf_zalesak(N[0]), rigid_rotation(N, T, **kw)
    elif kind == "C2":
        sdf, u = sdf_sphere([0.35 * n for n in N], 0.15 * N[0]), enright(N, T, **kw)
    elif kind == "C3":
        sdf, u = sdf_dambreak(N), tgv(N, T, **kw)
    elif kind == "C4":
        sdf, u = sdf_sphere([N[0] / 2, N[1] / 2, N[2] / 4], N[0] / 8, inside_dark=False), tgv(N, T, **kw)
    elif kind == "C5":
        sdf, u = sdf_sloshing(N), tgv(N, T, **kw)
    else:
        raise KeyError(kind)
    if vel == "enright" and len(N) == 3 and kind != "C2":
        u = enright(N, T, **kw)
    elif vel == "tgv" and kind in ("C2",):
        u = tgv(N, T, **kw)
    cfg.update(sdf=sdf, u=u, dtype=dtype, kind=kind)
    return cfg
