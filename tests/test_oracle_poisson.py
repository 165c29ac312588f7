"""Self-derived properties of the oracle's pressure projection (psolver!, inproject!, myproject!,
/root/reference/src/flow.jl:300-347, on WaterLily's Poisson struct).  The reference holds no known-answer test for these
(parity unpinned, oracle/oracle_poisson.hpp header); the checks here are independent of the restatement: a dense matrix assembled
in numpy from the definition of the operator, a dense least-squares solve, the discrete divergence after the projection.  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import pyoracle as O


def make_L(Ng, perdir, dtype, seed=0, lam_rho=1e-2):
    """μ₀ as MPFMomStep! leaves it: 1/ρ(f̄) on the faces of a random f, then BC!(μ₀, 0, false, perdir) (flow.jl:73,80,254-259)."""
    rng = np.random.default_rng(seed)
    D = len(Ng)
    f = np.asfortranarray(rng.uniform(0, 1, Ng).astype(dtype))
    O.BCf(f, perdir)
    mu0 = np.asfortranarray(np.ones(Ng + (D,), dtype))
    O.updateL(mu0, f, lam_rho, perdir)
    return mu0


def dense_operator(L, perdir):
    """A[I,J] of mult(I,L,D,x) over inside cells, assembled from the definition (periodic neighbours wrapped)."""
    Ng = L.shape[:-1]
    D = len(Ng)
    N = [n - 2 for n in Ng]
    idx = {I: k for k, I in enumerate(itertools.product(*[range(2, n) for n in Ng][::-1]))}  # any fixed order
    A = np.zeros((len(idx), len(idx)))
    for I, k in idx.items():
        I = I[::-1]
        k = idx[I[::-1]]
        for i in range(D):
            lo = L[tuple(c - 1 for c in I) + (i,)]
            Iu = list(I); Iu[i] += 1
            up = L[tuple(c - 1 for c in Iu) + (i,)]
            A[k, k] -= lo + up
            for s, w in ((-1, lo), (+1, up)):
                J = list(I); J[i] += s
                if J[i] < 2 or J[i] > Ng[i] - 1:
                    if (i + 1) in perdir:
                        J[i] = (J[i] - 2) % N[i] + 2
                    else:
                        assert w == 0  # BC!(μ₀,0): no coupling through a wall
                        continue
                A[k, idx[tuple(J)[::-1]]] += w
    return A, idx


def to_vec(a, idx):
    v = np.zeros(len(idx))
    for I, k in idx.items():
        v[k] = a[tuple(c - 1 for c in I[::-1])]
    return v


@pytest.mark.parametrize("Ng,perdir", [((9, 8), ()), ((9, 8), (1,)), ((8, 8), (1, 2)), ((7, 6, 6), ()), ((7, 6, 6), (2,)), ((6, 6, 6), (1, 2, 3))])
def test_operator_is_the_dense_matrix(Ng, perdir):
    L = make_L(Ng, perdir, np.float64, seed=1)
    A, idx = dense_operator(L, perdir)
    assert np.allclose(A, A.T, atol=1e-14)                          # symmetric
    assert np.abs(A.sum(axis=1)).max() < 1e-12                      # A·1 = 0 (pure Neumann / periodic)
    assert np.linalg.eigvalsh(A).max() < 1e-10                      # negative semi-definite
    rng = np.random.default_rng(2)
    x = np.asfortranarray(rng.standard_normal(Ng))
    p = O.Poisson(x, L, O.zeros(Ng, np.float64), perdir)
    z = O.pois_mult(p, x)
    assert np.allclose(to_vec(z, idx), A @ to_vec(x, idx), rtol=1e-12, atol=1e-12)
    assert np.allclose(to_vec(p.D, idx), np.diag(A), rtol=1e-14)
    inside = tuple(slice(1, -1) for _ in Ng)
    assert np.allclose(p.iD[inside] * p.D[inside], 1.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", [((12, 10), ()), ((10, 10), (1,)), ((8, 7, 6), ()), ((8, 8, 6), (1, 2))])
def test_psolver_solves_the_system(Ng, perdir, dtype):
    L = make_L(Ng, perdir, dtype, seed=3)
    A, idx = dense_operator(L.astype(np.float64), perdir)
    rng = np.random.default_rng(4)
    z = O.zeros(Ng, dtype)
    inside = tuple(slice(1, -1) for _ in Ng)
    b = rng.standard_normal(tuple(n - 2 for n in Ng))
    z[inside] = (b - b.mean()).astype(dtype)                        # compatible source
    x = O.zeros(Ng, dtype)
    p = O.Poisson(x, L, z, perdir)
    zv = to_vec(z.astype(np.float64), idx)                          # the solver uses z as scratch (flow.jl:312,316)
    n, r2 = O.psolver(p)
    eps = np.finfo(dtype).eps
    assert 0 < n < 2000 and r2 <= 50 * eps
    xr = np.linalg.lstsq(A, zv, rcond=None)[0]
    xv = to_vec(x.astype(np.float64), idx)
    d = (xv - xv.mean()) - (xr - xr.mean())
    assert np.abs(d).max() <= (2e-5 if dtype == np.float64 else 2e-1) * max(1.0, np.abs(xr).max())
    res = A @ xv - zv
    assert (res - res.mean()) @ (res - res.mean()) <= 4 * 50 * eps
    for j in perdir:                                                # perBC!(x) at the end: ghost planes wrap
        a = np.moveaxis(x, j - 1, 0)
        assert np.array_equal(a[0], a[-2]) and np.array_equal(a[-1], a[1])


def test_loop_condition_and_iteration_cap():
    Ng = (10, 9)
    L = make_L(Ng, (), np.float64, seed=5)
    z, x = O.zeros(Ng, np.float64), O.zeros(Ng, np.float64)
    p = O.Poisson(x, L, z, ())
    assert O.psolver(p) == (0, 0.0)                                 # r₂ = 0: no iteration, x untouched
    assert not x.any()
    rng = np.random.default_rng(6)
    b = rng.standard_normal((8, 7)); z[1:-1, 1:-1] = b - b.mean()
    n, r2 = O.psolver(p, itmx=3)
    assert n == 3 and r2 > 50 * np.finfo(np.float64).eps             # capped
    x[...] = 0
    z[1:-1, 1:-1] *= 1e-9                                           # tol/4 < r₂ <= tol: exactly one iteration (flow.jl:309)
    r0 = float((z[1:-1, 1:-1] ** 2).sum())
    z[1:-1, 1:-1] *= np.sqrt(0.5 * 50 * np.finfo(np.float64).eps / r0)
    n, _ = O.psolver(p)
    assert n == 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", [((14, 12), ()), ((12, 12), (1, 2)), ((10, 9, 8), ()), ((10, 8, 8), (2, 3))])
def test_myproject_removes_the_divergence(Ng, perdir, dtype):
    D = len(Ng)
    L = make_L(Ng, perdir, dtype, seed=7)
    rng = np.random.default_rng(8)
    u = np.asfortranarray(rng.standard_normal(Ng + (D,)).astype(dtype))
    O.BC(u, (0.0,) * D, False, perdir)
    x, z = O.zeros(Ng, dtype), O.zeros(Ng, dtype)
    p = O.Poisson(x, L, z, perdir)
    dt = 0.37
    n, r2 = O.myproject(u, p, dt)
    assert n > 0
    O.BC(u, (0.0,) * D, False, perdir)
    inside = tuple(slice(1, -1) for _ in Ng)
    div = np.zeros(tuple(n - 2 for n in Ng))
    for i in range(D):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(D)) + (i,)
        div += u[hi].astype(np.float64) - u[inside + (i,)].astype(np.float64)
    eps = np.finfo(dtype).eps
    assert (div ** 2).sum() <= 8 * 50 * eps                         # ‖∇·u‖² at the solver tolerance
    # x is handed back un-scaled (pressure): running the projection again changes nothing beyond round-off
    u2 = u.copy(order="F")
    O.myproject(u2, p, dt)
    assert np.abs(u2 - u).max() <= (1e-6 if dtype == np.float64 else 3e-3)


def test_periodic_shift_invariance():
    Ng, perdir = (10, 10), (1, 2)
    L = np.asfortranarray(np.ones(Ng + (2,)))
    rng = np.random.default_rng(9)
    b = rng.standard_normal((8, 8)); b -= b.mean()
    sols = []
    for s in (0, 3):
        z, x = O.zeros(Ng, np.float64), O.zeros(Ng, np.float64)
        z[1:-1, 1:-1] = np.roll(b, s, axis=0)
        p = O.Poisson(x, L, z, perdir)
        O.psolver(p)
        sols.append(np.roll(x[1:-1, 1:-1], -s, axis=0))
    assert np.abs(sols[0] - sols[1]).max() < 1e-6


# ---- WaterLily.MultiLevelPoisson (inproject!'s second method, flow.jl:343-347: solver!(b;tol=1e-4,itmx=200)) ---------------------------------
def test_multilevel_levels_and_restricted_coefficients():
    """levels: N+2 halves while divisible (even and > 4); restrictL!: 0.5·Σ of the fine faces that tile the coarse face, then BC!(·,0)."""
    Ng, perdir = (34, 18), ()
    L = make_L(Ng, perdir, np.float64, seed=11)
    ml = O.MultiLevelPoisson(O.zeros(Ng, np.float64), L, O.zeros(Ng, np.float64), perdir)
    shapes = [ml.level(l, "x").shape for l in range(ml.levels)]
    assert shapes == [(34, 18), (18, 10), (10, 6), (6, 4)]          # (10, 6) is still divisible (even and > 4); (6, 4) is not
    Lc = ml.level(1, "L")
    # coarse cell I (1-based) covers fine cells 2I-2, 2I-1: the x-face of coarse (5,4) = fine x-faces (8,6) and (8,7)
    I = (5, 4)
    fx = 0.5 * (L[2 * I[0] - 3, 2 * I[1] - 3, 0] + L[2 * I[0] - 3, 2 * I[1] - 2, 0])
    fy = 0.5 * (L[2 * I[0] - 3, 2 * I[1] - 3, 1] + L[2 * I[0] - 2, 2 * I[1] - 3, 1])
    assert Lc[I[0] - 1, I[1] - 1, 0] == fx and Lc[I[0] - 1, I[1] - 1, 1] == fy
    assert not Lc[1, :, 0].any() and not Lc[-1, :, 0].any()         # BC!(L,0): no coupling through the walls (normal faces)
    A, _ = dense_operator(np.asfortranarray(Lc.copy()), perdir)
    assert np.abs(A.sum(axis=1)).max() < 1e-12 and np.allclose(A, A.T)
    Dc = ml.level(1, "D")
    assert np.allclose(np.diag(A), to_vec(Dc, dense_operator(np.asfortranarray(Lc.copy()), perdir)[1]))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", [((18, 18), ()), ((34, 18), (1,)), ((18, 10, 10), ()), ((18, 18, 10), (1, 2))])
def test_multilevel_solver_solves_the_system(Ng, perdir, dtype):
    L = make_L(Ng, perdir, dtype, seed=12, lam_rho=1e-1)
    A, idx = dense_operator(L.astype(np.float64), perdir)
    rng = np.random.default_rng(13)
    z, x = O.zeros(Ng, dtype), O.zeros(Ng, dtype)
    inside = tuple(slice(1, -1) for _ in Ng)
    b = rng.standard_normal(tuple(n - 2 for n in Ng))
    z[inside] = (b - b.mean()).astype(dtype)
    zv = to_vec(z.astype(np.float64), idx)
    ml = O.MultiLevelPoisson(x, L, z, perdir)
    assert ml.levels >= 2
    ml.residual()
    r0 = float((ml.level(0, "r").astype(np.float64) ** 2).sum())
    ml.vcycle(); ml.smooth(0)
    r1 = float((ml.level(0, "r").astype(np.float64) ** 2).sum())
    assert r1 < 0.2 * r0                                            # one V-cycle + smoothing contracts the residual
    z[inside] = (b - b.mean()).astype(dtype)                        # pcg! used z as scratch: restore the source for solver!'s residual!
    n, r2 = ml.solver(tol=1e-4 if dtype == np.float32 else 1e-12, itmx=64)
    assert 0 < n < 64 and r2 < (1e-4 if dtype == np.float32 else 1e-12)
    xv = to_vec(x.astype(np.float64), idx)
    res = A @ xv - zv
    assert (res - res.mean()) @ (res - res.mean()) <= (4e-4 if dtype == np.float32 else 4e-12)
    xr = np.linalg.lstsq(A, zv, rcond=None)[0]
    d = (xv - xv.mean()) - (xr - xr.mean())
    assert np.abs(d).max() <= (3e-1 if dtype == np.float32 else 1e-4) * max(1.0, np.abs(xr).max())
    for j in perdir:
        a = np.moveaxis(x, j - 1, 0)
        assert np.array_equal(a[0], a[-2]) and np.array_equal(a[-1], a[1])


def test_multilevel_update_follows_a_changed_L():
    Ng = (18, 18)
    L = make_L(Ng, (), np.float64, seed=14)
    ml = O.MultiLevelPoisson(O.zeros(Ng, np.float64), L, O.zeros(Ng, np.float64), ())
    L2 = make_L(Ng, (), np.float64, seed=15)
    ref = O.MultiLevelPoisson(O.zeros(Ng, np.float64), L2, O.zeros(Ng, np.float64), ())
    L[...] = L2
    ml.update()
    for l in range(ml.levels):
        for name in ("L", "D", "iD"):
            assert np.array_equal(ml.level(l, name), ref.level(l, name))


@pytest.mark.parametrize("Ng,perdir", [((18, 18), ()), ((18, 10, 10), (2,))])
def test_multilevel_myproject_removes_the_divergence(Ng, perdir):
    D = len(Ng)
    dtype = np.float64
    L = make_L(Ng, perdir, dtype, seed=16, lam_rho=1e-1)
    rng = np.random.default_rng(17)
    u = np.asfortranarray(rng.standard_normal(Ng + (D,)).astype(dtype))
    O.BC(u, (0.0,) * D, False, perdir)
    x, z = O.zeros(Ng, dtype), O.zeros(Ng, dtype)
    ml = O.MultiLevelPoisson(x, L, z, perdir)
    n, r2 = O.ml_myproject(u, ml, 0.37)
    assert 0 < n <= 200 and r2 < 1e-4
    O.BC(u, (0.0,) * D, False, perdir)                              # flow.jl:82
    inside = tuple(slice(1, -1) for _ in Ng)
    div = np.zeros(tuple(n - 2 for n in Ng))
    for i in range(D):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(D)) + (i,)
        div += u[hi] - u[inside + (i,)]
    assert (div ** 2).sum() <= 4e-4                                 # ‖∇·u‖² at solver!'s tol = 1e-4 (flow.jl:346)
