"""Pins the oracle's explicit-forcing functions against the known answers the reference holds for them
(/root/reference/test/maintests.jl:65-70 getμ, :182-190 getPopinetHeight / getCurvature) and checks the unpinned rest
(visc!, surfTen!, updateU!, updateL!) through self-derived properties.  CPU only."""
import numpy as np
import pytest

from oracle import pyoracle as O

TS = [np.float32, np.float64]


def approx(T):
    return dict(rel=float(np.sqrt(np.finfo(T).eps)), abs=0.0)  # Julia's `≈`


@pytest.mark.parametrize("T", TS)
def test_getmu_kat(T):  # maintests.jl:65-70
    fFace = O.zeros((3, 3, 2), T)
    fFace[2, 1, 0] = 0.1; fFace[2, 2, 0] = 0.2; fFace[1, 2, 1] = 0.3; fFace[2, 2, 1] = 0.4
    Iur = (3, 3)
    assert O.getmu(1, 1, Iur, fFace, 0.1, 0.2, 1) == pytest.approx(0.02, **approx(T))
    a, b = O.getmu(1, 2, Iur, fFace, 0.1, 0.2, 0.2), O.getmu(2, 1, Iur, fFace, 0.1, 0.2, 0.2)
    assert a == b
    assert a == pytest.approx(0.028, **approx(T))


def _st_f(T):
    return np.asfortranarray(np.array([[0.3, 0.2, 0.1, 0.1, 0.2, 0.3, 0.0, 0.0],
                                       [1.0, 1.0, 0.6, 0.5, 0.3, 0.2, 0.0, 0.0],
                                       [0.0, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0, 0.0]], dtype=T))


def test_popinet_height_and_curvature_kat():  # maintests.jl:182-190 (Float64 literals)
    f = _st_f(np.float64)
    assert O.getPopinetHeight((1, 5), f, 2) == pytest.approx(-0.3, abs=1e-15)
    assert O.getPopinetHeight((2, 5), f, 2) == pytest.approx(-0.9, abs=1e-15)
    assert O.getPopinetHeight((3, 5), f, 2) == pytest.approx(-1.4, abs=1e-15)
    assert O.getCurvature((2, 5), f, 2) == pytest.approx(0.0672718547928328, rel=1e-12)
    f32 = _st_f(np.float32)
    assert O.getCurvature((2, 5), f32, 2) == pytest.approx(0.0672718547928328, rel=1e-5)


@pytest.mark.parametrize("T", TS)
@pytest.mark.parametrize("D", [2, 3])
def test_curvature_of_a_sphere(T, D):
    """Height-function curvature of a circle / sphere of radius R resolved with ~10 cells: -(D-1)/R within a few percent
    (self-derived; sign: the light phase f=0 is outside, heights are measured towards it)."""
    R, n = 10.3, 36
    Ng = (n + 2,) * D
    c = np.full(D, n / 2 + 0.37)
    f = O.zeros(Ng, T); al = O.zeros(Ng, T); nh = O.zeros(Ng + (D,), T)
    O.applyVOF(f, al, nh, lambda x: (np.sqrt(((x - c) ** 2).sum(-1)) - R))
    O.BCf(f, ())
    # the cell on the +x axis of the sphere that holds the interface: x centre index = c + R
    I = [int(round(c[0] + R + 1.5))] + [int(round(c[k] + 1.5)) for k in range(1, D)]
    while not (0 < f[tuple(i - 1 for i in I)] < 1):
        I[0] -= 1
    fI = f[tuple(i - 1 for i in I)]
    assert 0 < fI < 1
    # which phase is inside?  applyVOF! gives f=1 where sdf<0 (inside the sphere): the dark phase is inside, the normal points outwards (+x)
    k = O.getCurvature(tuple(I), f, 1)
    assert abs(k) == pytest.approx((D - 1) / R, rel=0.05)


@pytest.mark.parametrize("T", TS)
@pytest.mark.parametrize("N,perdir", [((12, 10), ()), ((12, 10), (1,)), ((8, 7, 6), ()), ((8, 7, 6), (2, 3))])
def test_visc_conserves_momentum_and_vanishes_for_rigid_motion(T, N, perdir):
    """visc! is a flux-difference form: with periodic BCs Σ r[inside,i] = 0 to round-off; a uniform translation gives r ≡ 0 exactly."""
    rng = np.random.default_rng(3)
    D = len(N)
    Ng = tuple(n + 2 for n in N)
    f = O.zeros(Ng, T); f[...] = rng.uniform(0, 1, Ng).astype(T); O.BCf(f, perdir)
    u = O.zeros(Ng + (D,), T); u[...] = 0.7
    r = O.zeros(Ng + (D,), T); r[...] = 5.0
    Phi = O.zeros(Ng, T); al = O.zeros(Ng, T); nh = O.zeros(Ng + (D,), T); fb = O.zeros(Ng, T)
    O.viscSurfTenrhou(r, u, Phi, f, al, nh, fb, 0.1, 0.3, 1e-2, None, perdir)
    assert np.all(r == 0)
    if len(perdir) == D or True:
        u[...] = rng.standard_normal(Ng + (D,)).astype(T)
        pd = tuple(range(1, D + 1))
        O.BCf(f, pd); O.BC(u, (0,) * D, False, pd)
        O.viscSurfTenrhou(r, u, Phi, f, al, nh, fb, 0.1, 0.3, 1e-2, None, pd)
        sl = tuple([slice(1, -1)] * D)
        for i in range(D):
            s = float(np.sum(r[sl + (i,)].astype(np.float64)))
            assert abs(s) <= (1e-3 if T == np.float32 else 1e-11) * np.abs(r[sl + (i,)]).sum()


@pytest.mark.parametrize("T", TS)
def test_visc_uniform_fluid_is_the_vector_laplacian_plus_grad_div(T):
    """f ≡ 1 (single phase, μ constant): r_i = μ Σ_j ∂_j(∂_j u_i + ∂_i u_j) with second differences (self-derived, periodic box)."""
    rng = np.random.default_rng(5)
    N, D = (9, 8, 7), 3
    Ng = tuple(n + 2 for n in N); pd = (1, 2, 3)
    f = O.zeros(Ng, T); f[...] = 1
    u = O.zeros(Ng + (D,), T); u[...] = rng.standard_normal(Ng + (D,)).astype(T); O.BC(u, (0, 0, 0), False, pd)
    r = O.zeros(Ng + (D,), T); Phi = O.zeros(Ng, T); al = O.zeros(Ng, T); nh = O.zeros(Ng + (D,), T); fb = O.zeros(Ng, T)
    mu = 0.25
    O.viscSurfTenrhou(r, u, Phi, f, al, nh, fb, 0.1, mu, 1e-2, None, pd)
    ud = u.astype(np.float64)
    sl = (slice(1, -1),) * 3
    for i in range(D):
        ref = np.zeros(N)
        for j in range(D):
            def F(shift):  # viscous flux at the lower j-face of momentum cell I + shift*δj
                a = np.roll(ud[..., i], -shift, axis=j) - np.roll(ud[..., i], -shift + 1, axis=j)
                b = np.roll(ud[..., j], -shift, axis=j) - np.roll(np.roll(ud[..., j], -shift, axis=j), 1, axis=i)
                return mu * (a + b)
            # rolls wrap through the ghost layer; evaluate on the interior where no wrap is involved except via BC!-periodic ghosts
            ref += (F(1) - F(0))[sl]
        got = r[sl + (i,)].astype(np.float64)
        inner = (slice(1, -1),) * 3  # stay one cell away from the ghost layer (np.roll wraps over the ghosts, BC! over the interior)
        assert np.abs(got[inner] - ref[inner]).max() <= (2e-5 if T == np.float32 else 1e-13)


@pytest.mark.parametrize("T", TS)
def test_update_u_and_update_l(T):
    rng = np.random.default_rng(7)
    N, D = (7, 6, 5), 3
    Ng = tuple(n + 2 for n in N)
    f = O.zeros(Ng, T); f[...] = rng.uniform(0, 1, Ng).astype(T); O.BCf(f, ())
    lr, dt, w = 1e-2, 0.3, 0.5
    mk = lambda: np.asfortranarray(rng.standard_normal(Ng + (D,)).astype(T))
    u, ru, ru0, fo = mk(), mk(), mk(), mk()
    ru_in, fo_in, u_in = ru.copy(), fo.copy(), u.copy()
    g = (0.0, -9.81, 0.5)
    O.updateU(u, ru, ru0, fo, dt, f, lr, g, w)
    a = T(1) / T(w) - T(1)
    ru_ref = (a * ru0 + ru_in + fo_in * T(dt)) * T(w)
    assert np.array_equal(ru, ru_ref)
    rho = O.zeros(Ng + (D,), T)
    for d in range(D):
        rho[..., d] = T(lr) + (T(1) - T(lr)) * ((f + np.roll(f, 1, axis=d)) / T(2))
    sl = (slice(1, -1),) * 3
    for d in range(D):
        exp_in = ru[sl + (d,)] / rho[sl + (d,)] + T(dt * w) * T(g[d])
        assert np.abs(u[sl + (d,)] - exp_in).max() <= 4 * np.finfo(T).eps * np.abs(exp_in).max()
        assert np.all(fo[..., d] == T(g[d]))
        # outside inside(f) only the gravity increment is applied
        ghost = np.ones(Ng, bool); ghost[sl] = False
        assert np.allclose(u[..., d][ghost], u_in[..., d][ghost] + T(dt * w) * T(g[d]), rtol=1e-6)
    mu0 = O.zeros(Ng + (D,), T); mu0[...] = 1
    O.updateL(mu0, f, lr, ())
    for d in range(D):
        sd = tuple(slice(2, -1) if k == d else slice(1, -1) for k in range(D))  # plane 2 of the normal component is a Dirichlet plane of BC!
        assert np.abs(mu0[sd + (d,)] - 1 / rho[sd + (d,)]).max() <= 4 * np.finfo(T).eps * (1 / lr)
    assert np.all(mu0[0, :, :, 0] == 0) and np.all(mu0[1, :, :, 0] == 0) and np.all(mu0[-1, :, :, 0] == 0)
    assert np.array_equal(mu0[0, :, :, 1], mu0[1, :, :, 1])  # Neumann copy of a tangential component
