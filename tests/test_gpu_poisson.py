"""Parity of the sm_100a pressure projection (update!, psolver!, myproject!; include/ifadv.h) against the CPU
oracle (oracle/oracle_poisson.hpp, self-checked against a dense operator in tests/test_oracle_poisson.py).
The per-cell arithmetic is IEEE-exact and follows the reference expression by expression, so set_diag! is compared BITWISE; the
conjugate-gradient recurrence goes through dot products whose summation order differs (as it does between the reference's own CPU and
GPU back ends), so a fixed number of iterations is compared to a few ulps of growth and the converged solve through the solver's
own tolerance: residual, divergence, iteration count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as O  # noqa: E402
from tests.helpers import alloc_cmom, dirO_for, make_state, oracle_mom_step_forcing  # noqa: E402
from tests.test_oracle_poisson import make_L  # noqa: E402


@pytest.fixture(scope="module")
def ia():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import interfaceadvection.jl_b200 as ia

    return ia


GRIDS = [((26, 18), ()), ((26, 18), (1,)), ((34, 34), (1, 2)), ((22, 16, 14), ()), ((20, 18, 14), (1, 2)), ((18, 18, 18), (1, 2, 3)),
         ((150, 11, 9), (3,))]


def _source(Ng, T, seed):
    rng = np.random.default_rng(seed)
    z = O.zeros(Ng, T)
    b = rng.standard_normal(tuple(n - 2 for n in Ng))
    z[tuple(slice(1, -1) for _ in Ng)] = (b - b.mean()).astype(T)
    return z


def _pair(ia, Ng, perdir, T, seed=3, x0=None):
    """The same Poisson problem on the oracle and on the device."""
    L = make_L(Ng, perdir, T, seed=seed)
    z = _source(Ng, T, seed + 1)
    x = O.zeros(Ng, T) if x0 is None else x0.copy(order="F")
    po = O.Poisson(x, L, z, perdir)
    pd = ia.Poisson(ia.from_numpy(x), ia.from_numpy(L), ia.from_numpy(z), perdir)
    return po, pd


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", GRIDS)
def test_update_matches_oracle_bitwise(ia, Ng, perdir, T):
    po, pd = _pair(ia, Ng, perdir, T)
    assert np.array_equal(ia.to_numpy(pd.D), po.D) and np.array_equal(ia.to_numpy(pd.iD), po.iD)
    L0 = O.zeros(Ng + (len(Ng),), T)                                # a cell cut off by zero coefficients: iD = 0 (abs2(D) < 2eps)
    pz = ia.Poisson(ia.from_numpy(O.zeros(Ng, T)), ia.from_numpy(L0), ia.from_numpy(O.zeros(Ng, T)), perdir)
    assert not ia.to_numpy(pz.iD).any()


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", GRIDS)
def test_fixed_iterations_match_oracle(ia, Ng, perdir, T):
    """Five iterations, far from convergence, from a non-zero initial guess: x, r, ϵ, z against the oracle."""
    rng = np.random.default_rng(12)
    x0 = np.asfortranarray(rng.standard_normal(Ng).astype(T))
    po, pd = _pair(ia, Ng, perdir, T, x0=x0)
    no, r2o = O.psolver(po, itmx=5)
    nd = ia.psolver(pd, itmx=5)
    assert no == 5 and nd == 5
    rel = 1e-11 if T == np.float64 else 2e-4
    assert pd.r2[-1] == pytest.approx(r2o, rel=rel)
    for name in ("x", "r", "eps", "z"):
        ref = getattr(po, name)
        got = ia.to_numpy(getattr(pd, name))
        assert np.abs(got - ref).max() <= rel * max(1.0, np.abs(ref).max()), name
    for j in perdir:                                                # perBC!(x) at the end
        a = np.moveaxis(ia.to_numpy(pd.x), j - 1, 0)
        assert np.array_equal(a[0], a[-2]) and np.array_equal(a[-1], a[1])
    if not perdir:                                                  # ghost entries of x are never written without a periodic direction
        g = ia.to_numpy(pd.x)
        assert np.array_equal(g[0], x0[0]) and np.array_equal(g[-1], x0[-1])


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", GRIDS)
def test_converged_solve_matches_oracle(ia, Ng, perdir, T):
    po, pd = _pair(ia, Ng, perdir, T, seed=21)
    z0 = po.z.copy(order="F")
    no, r2o = O.psolver(po)
    nd = ia.psolver(pd)
    eps = np.finfo(T).eps
    assert pd.r2[-1] <= 50 * eps and 0 < nd < 2000
    assert abs(nd - no) <= max(2, no // 10)                          # the loop stops on r₂ <= 50eps: round-off can move it by an iteration
    x = ia.to_numpy(pd.x)
    # the residual of the device solution, evaluated by the oracle's operator
    chk = O.Poisson(x.copy(order="F"), po.L, z0.copy(order="F"), perdir)
    O.pois_residual(chk)
    assert float((chk.r.astype(np.float64) ** 2).sum()) <= 4 * 50 * eps
    scale = max(1.0, np.abs(po.x).max())
    assert np.abs(x - po.x).max() <= (1e-6 if T == np.float64 else 5e-2) * scale


def test_loop_condition_and_iteration_cap(ia):
    Ng, T = (20, 14, 12), np.float64
    po, pd = _pair(ia, Ng, (), T, seed=30)
    pd.z.zero_()
    assert ia.psolver(pd) == 0 and pd.r2[-1] == 0.0                 # r₂ = 0: no iteration
    assert not ia.to_numpy(pd.x).any()
    z = _source(Ng, T, 31)
    pd.z.copy_(ia.from_numpy(z))
    assert ia.psolver(pd, itmx=3) == 3 and pd.r2[-1] > 50 * np.finfo(T).eps
    pd.x.zero_()
    z *= np.sqrt(0.5 * 50 * np.finfo(T).eps / float((z ** 2).sum()))  # tol/4 < r₂ <= tol: exactly one iteration (flow.jl:309)
    pd.z.copy_(ia.from_numpy(z))
    assert ia.psolver(pd) == 1
    pd.z.copy_(ia.from_numpy(_source(Ng, T, 32)))
    n1 = ia.psolver(pd, itmx=77)                                    # a cap that is not a multiple of the polling batch
    assert n1 == 77 or pd.r2[-1] <= 50 * np.finfo(T).eps
    pd.z.fill_(float("nan"))
    with pytest.raises(ia.IfadvError):
        ia.psolver(pd)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("Ng,perdir", GRIDS)
def test_myproject_matches_oracle(ia, Ng, perdir, T):
    D = len(Ng)
    L = make_L(Ng, perdir, T, seed=40)
    rng = np.random.default_rng(41)
    u = np.asfortranarray(rng.standard_normal(Ng + (D,)).astype(T))
    O.BC(u, (0.0,) * D, False, perdir)
    x0 = np.asfortranarray((0.1 * rng.standard_normal(Ng)).astype(T))   # the previous step's pressure as initial guess
    N = tuple(n - 2 for n in Ng)
    a = ia.Flow(N, (0.0,) * D, T=getattr(torch, np.dtype(T).name), dt=0.37, perdir=perdir)
    a.u.copy_(ia.from_numpy(u)); a.mu0.copy_(ia.from_numpy(L)); a.p.copy_(ia.from_numpy(x0))
    a.sigma.copy_(ia.from_numpy(np.asfortranarray(rng.standard_normal(Ng).astype(T))))  # garbage: inproject! overwrites it
    pd = ia.Poisson(a.p, a.mu0, a.sigma, perdir)
    pd.eps.fill_(3.0); pd.r.fill_(-2.0)
    uo, xo = u.copy(order="F"), x0.copy(order="F")
    po = O.Poisson(xo, L, O.zeros(Ng, T), perdir)
    no, _ = O.myproject(uo, po, T(0.5) * T(0.37))
    nd = ia.myproject(a, pd, 0.5)
    assert abs(nd - no) <= max(2, no // 10)
    ug, xg = ia.to_numpy(a.u), ia.to_numpy(a.p)
    tol = 1e-6 if T == np.float64 else 5e-2
    assert np.abs(ug - uo).max() <= tol * max(1.0, np.abs(uo).max())
    assert np.abs(xg - xo).max() <= tol * max(1.0, np.abs(xo).max())
    inside = tuple(slice(1, -1) for _ in Ng)
    ia.BC(a.u, (0.0,) * D, False, perdir)
    ug = ia.to_numpy(a.u)
    div = np.zeros(N)
    for i in range(D):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(D)) + (i,)
        div += ug[hi].astype(np.float64) - ug[inside + (i,)].astype(np.float64)
    assert (div ** 2).sum() <= 8 * 50 * np.finfo(T).eps              # ‖∇·u‖² at the solver tolerance
    # entries myproject! does not touch: faces outside inside(x)
    top = tuple(slice(-1, None) if d == 0 else slice(None) for d in range(D))
    assert np.array_equal(ia.to_numpy(a.u)[top], ug[top])


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir,shift", [((32, 24), "C1", (), None), ((24, 20, 16), "C3", (), None),
                                                  ((24, 16, 20), "C4", (1, 2), (0.37, 0.21, 0.13))])
def test_full_step_with_projection_matches_oracle(ia, T, N, kind, perdir, shift):
    """Two complete MPFMomStep! -- transport, viscosity, surface tension, gravity AND the reference's own pressure projection -- on the
    B200 kernels against the oracle doing the same (psolver=Poisson)."""
    st = make_state(N, kind, T, perdir=perdir, scale_u=0.5, shift=shift)
    D = st["D"]
    mu, lam_mu, eta = 0.02, 0.05, 0.05
    g = (0.0, -0.01, 0.0)[:D]
    dt = 0.4
    sim = ia.TwoPhaseSimulation(N, (0,) * D, float(N[0]), T=getattr(torch, np.dtype(T).name), lam_mu=lam_mu, lam_rho=st["lam_rho"], eta=eta,
                                nu=mu, g=g, perdir=perdir, dt=dt, psolver="Poisson")
    a, c, b = sim.flow, sim.intf, sim.pois
    c.f.copy_(ia.from_numpy(st["f"])); a.u.copy_(ia.from_numpy(st["u"])); a.dt[:] = [dt]
    fo, uo = st["f"].copy(order="F"), st["u"].copy(order="F")
    ao = alloc_cmom(st); ao["mu0"] = O.zeros(st["Ng"] + (D,), T); ao["mu0"][...] = 1
    po = O.Poisson(O.zeros(st["Ng"], T), ao["mu0"], ao["Phi"], perdir)
    tol = 1e-6 if T == np.float64 else 2e-2
    for n in range(2):
        dirO = dirO_for(n, D)
        oracle_mom_step_forcing(st, ao, fo, uo, dt, dirO, mu, lam_mu, eta, g, pois=po)
        ia.mom_step_forcing(a, c, dt, project=ia.project_with(b))
        a.dt.append(dt)
        assert np.abs(ia.to_numpy(c.f) - fo).max() <= tol, n
        assert np.abs(ia.to_numpy(a.u) - uo).max() <= tol * max(1.0, np.abs(uo).max()), n
        assert np.abs(ia.to_numpy(a.p) - po.x).max() <= tol * max(1.0, np.abs(po.x).max()), n
    assert len(b.n) == 4 and all(k > 0 for k in b.n)
    # after the corrector's projection the velocity is discretely solenoidal to the solver tolerance
    ug = ia.to_numpy(a.u)
    inside = tuple(slice(1, -1) for _ in st["Ng"])
    div = np.zeros(N)
    for i in range(D):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(D)) + (i,)
        div += ug[hi].astype(np.float64) - ug[inside + (i,)].astype(np.float64)
    assert (div ** 2).sum() <= 8 * 50 * np.finfo(T).eps


def test_projection_at_a_named_size(ia):
    """C2's grid (256³, Float32): density ratio 1000 across a sphere, random divergent velocity -> solenoidal to the solver tolerance."""
    N = (256, 256, 256)
    T = torch.float32
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 256.0, T=T, lam_rho=1e-3, dt=0.25, psolver="Poisson",
                                InterfaceSDF=lambda x: ((x - 128.0) ** 2).sum(-1).sqrt() - 64.0)
    a, c, b = sim.flow, sim.intf, sim.pois
    gen = torch.Generator(device="cuda").manual_seed(5)
    a.u.copy_(0.1 * torch.randn(a.u.shape, generator=gen, device="cuda", dtype=T).to(a.u.dtype))
    ia.BC(a.u, a.uBC, False, ())
    ia.updateL(a.mu0, c.f, c.lam_rho, (), fill_one=True)
    ia.update(b)
    n = ia.myproject(a, b, 1.0)
    ia.BC(a.u, a.uBC, False, ())
    assert 0 < n <= 2000
    u = a.u
    div = torch.zeros(N, device="cuda", dtype=torch.float64)
    for i in range(3):
        hi = tuple(slice(2, None) if d == i else slice(1, -1) for d in range(3)) + (i,)
        lo = (slice(1, -1),) * 3 + (i,)
        div += u[hi].double() - u[lo].double()
    r2 = float((div ** 2).sum())
    assert n == 2000 or r2 <= 8 * 50 * np.finfo(np.float32).eps, (n, r2, b.r2[-1])
