"""Pins the CPU oracle against every known-answer test the reference holds for this path
(/root/reference/test/maintests.jl; line numbers cited per test).  CPU only."""
import numpy as np
import pytest

from oracle import pyoracle as O

TS = [np.float32, np.float64]


def approx(T):
    # Julia's `≈` default: rtol = sqrt(eps(T))
    return dict(rel=float(np.sqrt(np.finfo(T).eps)), abs=0.0)


@pytest.mark.parametrize("T", TS)
def test_plic_intercept_kat(T):  # maintests.jl:28-37
    A = approx(T)
    t = T
    assert O.getIntercept([t(2 / 3), t(4 / 3), t(0)], t(5 / 12), T) == pytest.approx(8 / 9, **A)
    assert O.getIntercept([t(2 / 3), t(4 / 3)], t(5 / 12), T) == pytest.approx(8 / 9, **A)
    assert O.getIntercept([t(2 / 3), t(0), t(4 / 3)], t(5 / 12), T) == pytest.approx(8 / 9, **A)
    assert O.getIntercept([t(2 / 3), -t(4 / 3), t(0)], t(5 / 12), T) == pytest.approx(-4 / 9, **A)
    assert O.getIntercept([t(2 / 3), -t(4 / 3)], t(5 / 12), T) == pytest.approx(-4 / 9, **A)
    assert O.getIntercept([t(3), -t(4), t(0)], t(5 / 6), T) == pytest.approx(1, **A)
    assert O.getIntercept([t(3), -t(4)], t(5 / 6), T) == pytest.approx(1, **A)
    assert O.getIntercept([t(1 / 2), t(1 / 3), t(1)], t(7 / 12), T) == pytest.approx(1, **A)
    assert O.getIntercept([t(1), t(1), -t(1)], t(1 - 1 / 48), T) == pytest.approx(3 / 2, **A)
    # 2-D call ≈ 3-D call with a zero third component (:35)
    a2 = O.getIntercept([-t(7 / 6), t(4 / 9)], t(7 / 23), T)
    a3 = O.getIntercept([-t(7 / 6), t(4 / 9), t(0)], t(7 / 23), T)
    assert a2 == pytest.approx(a3, **A)


@pytest.mark.parametrize("T", TS)
def test_plic_volume_fraction_kat(T):  # maintests.jl:39-48
    A = approx(T)
    t = T
    assert O.getVolumeFraction([t(2 / 3), t(4 / 3), t(0)], t(8 / 9), T) == pytest.approx(5 / 12, **A)
    assert O.getVolumeFraction([t(2 / 3), t(4 / 3)], t(8 / 9), T) == pytest.approx(5 / 12, **A)
    assert O.getVolumeFraction([t(2 / 3), t(0), t(4 / 3)], t(8 / 9), T) == pytest.approx(5 / 12, **A)
    assert O.getVolumeFraction([t(2 / 3), -t(4 / 3), t(0)], -t(4 / 9), T) == pytest.approx(5 / 12, **A)
    assert O.getVolumeFraction([t(2 / 3), -t(4 / 3)], -t(4 / 9), T) == pytest.approx(5 / 12, **A)
    assert O.getVolumeFraction([t(3), -t(4), t(0)], t(1), T) == pytest.approx(5 / 6, **A)
    assert O.getVolumeFraction([t(3), -t(4)], t(1), T) == pytest.approx(5 / 6, **A)
    assert O.getVolumeFraction([t(1 / 2), t(1 / 3), t(1)], t(1), T) == pytest.approx(7 / 12, **A)
    assert O.getVolumeFraction([t(1), t(1), -t(1)], t(3 / 2), T) == pytest.approx(1 - 1 / 48, **A)
    v2 = O.getVolumeFraction([-t(7 / 6), t(4 / 9)], t(7 / 23), T)
    v3 = O.getVolumeFraction([-t(7 / 6), t(4 / 9), t(0)], t(7 / 23), T)
    assert v2 == pytest.approx(v3, **A)


@pytest.mark.parametrize("T", TS)
def test_plic_roundtrip(T):
    """f -> α -> f identity over random normals (property implied by the KAT pairs :28-44)."""
    rng = np.random.default_rng(20261017)
    tol = 2e-5 if T == np.float32 else 1e-12
    for D in (2, 3):
        for _ in range(400):
            n = rng.uniform(-1, 1, D).astype(T)
            if rng.random() < 0.2:
                n[rng.integers(D)] = 0
            if np.sum(np.abs(n)) == 0:
                continue
            g = T(rng.uniform(1e-3, 1 - 1e-3))
            a = O.getIntercept(list(n), g, T)
            assert abs(O.getVolumeFraction(list(n), T(a), T) - float(g)) < tol


def test_vofutil_kat():  # maintests.jl:55-63, 73-77
    f = O.zeros((3, 3), np.float64)
    f[1, 1] = 0.32
    f[1, 2] = 0.64
    # get3CellHeight(f,Ic,2) ≈ 0.96 through a Column normal is indirect; check getρ via u2ρu!/f2face!
    fF = O.zeros((3, 3, 2), np.float64)
    O.f2face(fF, f)
    lr = 0.7
    assert lr + (1 - lr) * f[1, 1] == pytest.approx(0.796)          # getρ(Ic,f,0.7)
    assert lr + (1 - lr) * fF[1, 1, 1] == pytest.approx(0.748)      # getρ(2,Ic,f,0.7)
    u = O.zeros((3, 3, 2), np.float64)
    u[...] = 1.0
    ru = O.zeros((3, 3, 2), np.float64)
    O.u2rhou(ru, u, f, lr)
    assert ru[1, 1, 1] == pytest.approx(0.748)
    # f2face! (:73-77)
    fC = O.zeros((4, 4), np.float64)
    fC[1:3, 1:3] = [[0.2, 0.6], [0.4, 0.8]]
    fF = O.zeros((4, 4, 2), np.float64)
    O.f2face(fF, fC)
    assert fF[2, 2, 0] == pytest.approx(0.7) and fF[2, 2, 0] == pytest.approx((fC[1, 2] + fC[2, 2]) / 2)
    assert fF[2, 2, 1] == pytest.approx(0.6) and fF[2, 2, 1] == pytest.approx((fC[2, 1] + fC[2, 2]) / 2)


def test_bc_kat():  # maintests.jl:82-129
    rng = np.random.default_rng(1)
    g = np.asfortranarray(rng.random((6, 6)))
    gN = g.copy(order="F"); O.BCf(gN)
    assert (gN[0, :] == gN[1, :]).all() and (gN[-1, :] == gN[-2, :]).all()
    assert (gN[:, 0] == gN[:, 1]).all() and (gN[:, -1] == gN[:, -2]).all()
    gP = g.copy(order="F"); O.BCf(gP, perdir=(1, 2))
    assert (gP[0, :] == gP[-2, :]).all() and (gP[-1, :] == gP[1, :]).all()
    assert (gP[:, 0] == gP[:, -2]).all() and (gP[:, -1] == gP[:, 1]).all()
    # BCf!(d,f) == BCv1D! (:96-101)
    g1 = g.copy(order="F"); O.BCv1D(g1, 1)
    assert (g1[0, 1:-1] == g[2, 1:-1]).all()
    assert (g1[-1, 1:-1] == g[-1, 1:-1]).all()
    assert (g1[:, 0] == g1[:, 1]).all() and (g1[:, -1] == g1[:, -2]).all()
    g1p = g.copy(order="F"); O.BCv1D(g1p, 1, perdir=(1,))
    assert (g1p[0, :] == g1p[-2, :]).all() and (g1p[-1, :] == g1p[1, :]).all()
    # BCv! (:105-115)
    u = np.asfortranarray(rng.random((6, 6, 2)))
    uN = u.copy(order="F"); O.BCv(uN)
    assert (uN[0, 1:-1, 0] == u[2, 1:-1, 0]).all()
    assert (uN[-1, 1:-1, 0] == u[-1, 1:-1, 0]).all()
    assert (uN[:, 0, 0] == uN[:, 1, 0]).all() and (uN[:, -1, 0] == uN[:, -2, 0]).all()
    assert (uN[1:-1, 0, 1] == u[1:-1, 2, 1]).all()
    assert (uN[1:-1, -1, 1] == u[1:-1, -1, 1]).all()
    assert (uN[0, :, 1] == uN[1, :, 1]).all() and (uN[-1, :, 1] == uN[-2, :, 1]).all()
    u1 = np.asfortranarray(u[:, :, 0].copy()); O.BCv1D(u1, 1)
    assert (u1 == uN[:, :, 0]).all()
    # BCVOF! (:119-129)
    f = np.asfortranarray(rng.random((6, 6))); al = np.asfortranarray(rng.random((6, 6))); nh = np.asfortranarray(rng.random((6, 6, 2)))
    a0, n0 = al.copy(), nh.copy()
    O.BCVOF(f, al, nh)
    assert (f[0, :] == f[1, :]).all() and (f[-1, :] == f[-2, :]).all()
    assert (al[0, :] == a0[0, :]).all() and (nh[0, :, :] == n0[0, :, :]).all()
    f = np.asfortranarray(rng.random((6, 6))); al = np.asfortranarray(rng.random((6, 6))); nh = np.asfortranarray(rng.random((6, 6, 2)))
    O.BCVOF(f, al, nh, perdir=(1, 2))
    assert (f[0, :] == f[-2, :]).all() and (al[0, :] == al[-2, :]).all() and (nh[0, :, :] == nh[-2, :, :]).all()


def _sdf_plane(x):
    return (-x[..., 0] - 3 * x[..., 1] + 4.5) / np.sqrt(x.dtype.type(10))


def test_applyvof_kat():  # maintests.jl:131-136 and :237-240
    f = O.zeros((4, 4), np.float64); al = O.zeros((4, 4), np.float64); nh = O.zeros((4, 4, 2), np.float64)
    O.applyVOF(f, al, nh, _sdf_plane)
    fRef = np.array([[0, 0, 0, 0], [0, 0, 2 / 3, 0], [0, 1 / 24, 23 / 24, 0], [0, 0, 0, 0]])
    assert np.allclose(f, fRef, rtol=1.5e-8, atol=1e-9)
    O.BCf(f)  # cVOF ctor (:237-240)
    fRef2 = np.array([[0, 0, 2 / 3, 2 / 3], [0, 0, 2 / 3, 2 / 3], [1 / 24, 1 / 24, 23 / 24, 23 / 24], [1 / 24, 1 / 24, 23 / 24, 23 / 24]])
    assert np.allclose(f, fRef2, rtol=1.5e-8, atol=1e-9)


def test_normals_kat():  # maintests.jl:140-155
    I = (2, 2)
    f = O.farr([[5 / 12, 1, 2 / 3], [1 / 4, 11 / 12, 1 / 12], [1 / 12, 1 / 3, 0]], np.float64)
    nh = O.zeros((3, 3, 2), np.float64)
    O.normal("WY", f, nh, I)
    assert nh[1, 1, 0] == pytest.approx(1.0) and nh[1, 1, 1] + 0.5 == pytest.approx(0.5)
    f = O.farr([[0, 1 / 3, 1], [1 / 12, 11 / 12, 1], [1, 1, 1]], np.float64)
    O.normal("WY", f, nh, I)
    assert nh[1, 1, 0] == pytest.approx(-2 / 3) and nh[1, 1, 1] == pytest.approx(-1.0)
    f = O.farr([[0, 0, 0], [0, 0.1, 0], [0, 0, 0]], np.float64)
    nh[...] = 0
    O.normal("WY", f, nh, I)
    assert nh[1, 1, 0] == 1 and nh[1, 1, 1] == 0
    nh[...] = 0
    O.normal("MYC", f, nh, I)
    assert nh[1, 1, 0] == 0.5 and nh[1, 1, 1] == 0.5


def test_vof_flux_kat():  # maintests.jl:159-179
    f = O.zeros((3, 3), np.float64); f[1, 1] = 0.32
    al = O.zeros((3, 3), np.float64); al[1, 1] = -0.2
    nh = O.zeros((3, 3, 2), np.float64); nh[1, 1, :] = [1, -1]
    ruf = O.zeros((3, 3, 2), np.float64); lr = 0.1
    ff = O.zeros((3, 3), np.float64)
    O.getVOFFlux_face(ff, f, al, nh, -0.4, 1, (2, 2), ruf, lr)
    O.getVOFFlux_face(ff, f, al, nh, 0.4, 1, (3, 2), ruf, lr)
    assert ff[1, 1] == pytest.approx(-0.24) and ff[2, 1] == pytest.approx(0.02)
    assert ruf[1, 1, 0] == pytest.approx(-0.256) and ruf[2, 1, 0] == pytest.approx(0.058)
    O.getVOFFlux_face(ff, f, al, nh, -0.4, 2, (2, 2), ruf, lr)
    O.getVOFFlux_face(ff, f, al, nh, 0.4, 2, (2, 3), ruf, lr)
    assert ff[1, 1] == pytest.approx(-0.02) and ff[1, 2] == pytest.approx(0.24)
    assert ruf[1, 1, 1] == pytest.approx(-0.058) and ruf[1, 2, 1] == pytest.approx(0.256)
