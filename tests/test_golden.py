"""Committed golden vectors (tests/golden/, generator: tests/golden/make_golden.py).

CPU part: the oracle reproduces the frozen fixtures bit for bit and the reference's own known-answer values
(reference_kats.json, transcribed from test/maintests.jl).  GPU part (-m gpu): the sm_100a path, called through the
C ABI, reproduces the same fixtures within the north-star tolerance (1e-12 Float64 / 1e-5 Float32)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.helpers import inside, make_state, oracle_mom_advect_step

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CMOM = sorted(glob.glob(os.path.join(HERE, "cmom_*.npz")))
VOF = sorted(glob.glob(os.path.join(HERE, "vof_*.npz")))
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def _state(z):
    T = z["f_in"].dtype.type
    st = dict(N=tuple(int(n) for n in z["N"]), D=len(z["N"]), Ng=tuple(int(n) + 2 for n in z["N"]), dtype=T,
              perdir=tuple(int(p) for p in z["perdir"]), uBC=tuple(float(a) for a in z["uBC"]) if "uBC" in z else None,
              f=np.asfortranarray(z["f_in"]), u=np.asfortranarray(z["u"]), lam_rho=float(z["lam_rho"]))
    return st


def test_fixture_set_is_complete():
    assert len(CMOM) >= 6 and len(VOF) >= 2


def test_reference_kats_on_oracle():
    K = json.load(open(os.path.join(HERE, "reference_kats.json")))
    for T in (np.float32, np.float64):
        rel = float(np.sqrt(np.finfo(T).eps))
        for c in K["getIntercept"]["cases"]:
            a = O.getIntercept([T(v) for v in c["n"]], T(c["g"]), T)
            assert a == pytest.approx(c["alpha"], rel=rel), c
        for c in K["getVolumeFraction"]["cases"]:
            v = O.getVolumeFraction([T(x) for x in c["n"]], T(c["alpha"]), T)
            assert v == pytest.approx(c["f"], rel=rel), c
    f = O.zeros((4, 4), np.float64); al = O.zeros((4, 4), np.float64); nh = O.zeros((4, 4, 2), np.float64)
    O.applyVOF(f, al, nh, lambda x: (-x[..., 0] - 3 * x[..., 1] + 4.5) / np.sqrt(10.0))
    assert np.allclose(f, np.array(K["applyVOF"]["fRef"]), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("path", CMOM, ids=[os.path.basename(p)[:-4] for p in CMOM])
def test_oracle_reproduces_cmom_golden(path):
    z = np.load(path)
    st = _state(z)
    # the inputs themselves are regenerated identically by the case builders
    st2 = make_state(st["N"], str(z["kind"]), st["dtype"], perdir=st["perdir"], uBC=st["uBC"])
    assert np.array_equal(st2["f"], st["f"]) and np.array_equal(st2["u"], st["u"])
    f = st["f"].copy(order="F")
    ru = oracle_mom_advect_step(st, f, st["u"], 1.0, tuple(int(d) for d in z["dirO"]))
    assert np.array_equal(f, z["f_out"])
    assert np.array_equal(inside(ru, st["D"]), inside(z["rhou_out"], st["D"]))


@pytest.mark.parametrize("path", VOF, ids=[os.path.basename(p)[:-4] for p in VOF])
def test_oracle_reproduces_vof_golden(path):
    z = np.load(path)
    st = _state(z)
    T, D, Ng = st["dtype"], st["D"], st["Ng"]
    f = st["f"].copy(order="F")
    zz = lambda *s: O.zeros(s, T)
    ff, al, nh, rhouf = zz(*Ng), zz(*Ng), zz(*Ng, D), zz(*Ng, D)
    cbar = np.zeros(Ng, dtype=np.int8, order="F")
    for _ in range(int(z["steps"])):
        O.advectVOF(f, ff, al, nh, st["u"], st["u"], 1.0, cbar, rhouf, st["lam_rho"], "WH", st["perdir"], tuple(int(d) for d in z["dirO"]))
    assert np.array_equal(f, z["f_out"])
    # discrete mass conservation of the frozen output (solenoidal velocity, closed / periodic box)
    assert abs(inside(z["f_out"], D).sum() - inside(z["f_in"], D).sum()) <= 1e-10 * inside(z["f_in"], D).sum()


# ------------------------------------------------------------------------------------------------ GPU: C ABI vs golden
@pytest.fixture(scope="module")
def ia():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import interfaceadvection.jl_b200 as ia

    return ia


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in CMOM if "_2d_" not in p], ids=[os.path.basename(p)[:-4] for p in CMOM if "_2d_" not in p])
def test_cuda_host_entry_matches_cmom_golden(ia, path):
    """ifadv_mom_advect_step_host (host buffers in, host buffers out) against the frozen vectors."""
    z = np.load(path)
    st = _state(z)
    T = st["dtype"]
    f_h = st["f"].copy(order="F"); ru_h = np.zeros_like(st["u"], order="F")
    ctx = ia.Context(st["Ng"], np.dtype(T).name, 0)
    rep = ia.Report()
    rc = ctx.mom_advect_step_host(f_h.ctypes.data, st["u"].ctypes.data, ru_h.ctypes.data, 1.0, st["lam_rho"], ia.LIMITERS["Koren"],
                                  ia.NORMAL_SCHEMES["WH"], st["uBC"], st["perdir"], tuple(int(d) for d in z["dirO"]), rep)
    assert rc == 0
    tol = TOL[np.dtype(T)]
    assert np.abs(f_h - z["f_out"]).max() <= tol
    assert np.abs(inside(ru_h, 3) - inside(z["rhou_out"], 3)).max() <= tol
    assert ctx.launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", CMOM, ids=[os.path.basename(p)[:-4] for p in CMOM])
def test_cuda_mirror_matches_cmom_golden(ia, path):
    """MPFMomStep (Python mirror of flow.jl:60-109, transport half) against the frozen vectors."""
    import torch

    z = np.load(path)
    st = _state(z)
    T, D = st["dtype"], st["D"]
    TT = getattr(torch, np.dtype(T).name)
    flow = ia.Flow(st["N"], st["uBC"], T=TT, dt=1.0, perdir=st["perdir"])
    intf = ia.cVOF(st["N"], T=TT, lam_rho=st["lam_rho"], perdir=st["perdir"])
    flow.u.copy_(ia.from_numpy(st["u"]))
    intf.f.copy_(ia.from_numpy(st["f"]))
    # the sweep order follows length(Δt) (flow.jl:163): pad Δt so that the mirror derives the fixture's dirO
    dirO = tuple(int(d) for d in z["dirO"])
    n = next(n for n in range(1, 8) if tuple((n + i) % D + 1 for i in range(1, D + 1)) == dirO)
    flow.dt = [1.0] * n
    ia.MPFMomStep(flow, None, intf, None, dt=1.0, check=True)
    torch.cuda.synchronize()
    tol = TOL[np.dtype(T)]
    assert np.abs(ia.to_numpy(intf.f) - z["f_out"]).max() <= tol
    assert np.abs(inside(ia.to_numpy(intf.rhou), D) - inside(z["rhou_out"], D)).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("path", VOF, ids=[os.path.basename(p)[:-4] for p in VOF])
def test_cuda_matches_vof_golden(ia, path):
    z = np.load(path)
    st = _state(z)
    T, D, Ng = st["dtype"], st["D"], st["Ng"]
    zz = lambda *s: ia.from_numpy(O.zeros(s, T))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    ff, al, nh, rhouf = zz(*Ng), zz(*Ng), zz(*Ng, D), zz(*Ng, D)
    cbar = ia.from_numpy(np.zeros(Ng, dtype=np.int8, order="F"))
    for _ in range(int(z["steps"])):
        ia.advectVOF(fd, ff, al, nh, ud, ud, 1.0, cbar, rhouf, st["lam_rho"], "WH", st["perdir"], tuple(int(d) for d in z["dirO"]))
    assert np.abs(ia.to_numpy(fd) - z["f_out"]).max() <= 10 * TOL[np.dtype(T)]  # three steps
