"""Parity of the sm_100a post-processing kernels (level-set redistancing, metrics; include/ifadv.h) against the
CPU oracle, which is pinned by the reference's own tests (tests/test_oracle_post.py ↔ test/maintests.jl:264-345).  Field results are
compared BITWISE in both precisions (the kernels are compiled IEEE-exact and follow the reference expression by expression); the
Float64-accumulated sums to round-off of the summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as O  # noqa: E402
from tests.helpers import inside, make_state  # noqa: E402


@pytest.fixture(scope="module")
def ia():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import interfaceadvection.jl_b200 as ia

    return ia


CASES = [((18, 16), "C1", ()), ((20, 14), "C1", (2,)), ((16, 12, 10), "C2", ()), ((14, 12, 16), "C4", (1, 2)), ((40, 9, 8), "C3", (1, 2, 3))]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir", CASES)
def test_redistancing_matches_oracle_bitwise(ia, T, N, kind, perdir):
    st = make_state(N, kind, T, perdir=perdir)
    D, Ng = st["D"], st["Ng"]
    rng = np.random.default_rng(2)
    phi = np.asfortranarray((2 * st["f"] - 1 + 0.05 * rng.standard_normal(Ng)).astype(T)); O.BCf(phi, perdir)
    pini = phi.copy(order="F")
    # computeL!
    L_o = O.zeros(Ng, T); O.computeL(L_o, phi, pini, perdir)
    L_d = ia.from_numpy(np.asfortranarray(rng.standard_normal(Ng).astype(T)))
    ia.computeL(L_d, ia.from_numpy(phi), ia.from_numpy(pini), perdir)
    assert np.array_equal(inside(ia.to_numpy(L_d), D), inside(L_o, D))
    # one stage
    p_o, p0 = phi.copy(order="F"), np.asfortranarray(phi + T(0.01))
    O.redistaningStage(p_o, p0, pini, L_o, 0.3, 0.75, perdir)
    p_d = ia.from_numpy(phi)
    ia.redistaningStage(p_d, ia.from_numpy(p0), ia.from_numpy(pini), L_d, 0.3, 0.75, perdir)
    assert np.array_equal(inside(ia.to_numpy(p_d), D), inside(p_o, D))
    # the whole reinitialisation: round(d/dτ) SSP-RK3 steps with BCf! after every stage
    p_o, p0_o, L_o = phi.copy(order="F"), O.zeros(Ng, T), O.zeros(Ng, T)
    O.redistaning(p_o, p0_o, pini, L_o, d=3, dtau=0.4, perdir=perdir)
    p_d, p0_d, L_d = ia.from_numpy(phi), ia.from_numpy(O.zeros(Ng, T)), ia.from_numpy(O.zeros(Ng, T))
    ia.context_for(p_d).redistance(0, p_d.data_ptr(), p0_d.data_ptr(), ia.from_numpy(pini).data_ptr(), L_d.data_ptr(), 3, 0.4, perdir)
    torch.cuda.synchronize()
    assert np.array_equal(ia.to_numpy(p_d), p_o)  # ghosts included (BCf!)
    assert np.isfinite(p_o).all()


def test_levelset_of_a_simulation_reproduces_the_reference_test(ia):
    """test/maintests.jl:294-318 through the mirror: LevelSet(sim) aliases f⁰ / α / fᶠ / σ, ϕ = 2f-1, and the reinitialised planar
    interface is a monotone, periodic signed-distance field close to -(x-8) near the interface."""
    sim = ia.TwoPhaseSimulation((16, 16), (0.0, 0.0), 16.0, T=torch.float64, InterfaceSDF=lambda x: x[..., 0] - 8, perdir=(2,))
    f0 = ia.to_numpy(sim.intf.f).copy()
    ls = ia.LevelSet(sim)
    assert ls.phi is sim.intf.f0 and ls.phi0 is sim.intf.alpha
    assert np.array_equal(ia.to_numpy(ls.phi), 2 * f0 - 1)
    ia.redistaning(ls, d=4, dtau=0.05, perdir=(2,))
    phi = ia.to_numpy(ls.phi)
    assert np.isfinite(phi).all()
    assert np.all(np.diff(phi[1:-1, 8]) <= 0)
    assert np.array_equal(phi[:, 0], phi[:, -2]) and np.array_equal(phi[:, -1], phi[:, 1])
    for ix in range(8, 12):
        assert phi[ix - 1, 8] == pytest.approx(-(ix - 1.5 - 8), abs=0.01)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind", [((24, 18), "C1"), ((20, 14, 12), "C2")])
def test_metrics_match_oracle(ia, T, N, kind):
    st = make_state(N, kind, T)
    D, Ng = st["D"], st["Ng"]
    rng = np.random.default_rng(9)
    u = np.asfortranarray(rng.standard_normal(Ng + (D,)).astype(T))
    U, g, W = (0.1, -0.2, 0.3)[:D], (0.0, -9.81, 0.5)[:D], (0.0, 3.0, 1.0)[:D]
    ke, pe, mom = O.metrics_sum(u, st["f"], st["lam_rho"], U, g, W)
    ke_c, pe_c, mom_c = ia.metrics(ia.from_numpy(u), ia.from_numpy(st["f"]), st["lam_rho"], U, g, W)
    rel = 1e-12 if T == np.float64 else 1e-12  # per-cell values are formed identically; only the Float64 summation order differs
    assert ke_c == pytest.approx(ke, rel=rel) and pe_c == pytest.approx(pe, rel=rel)
    assert mom_c == pytest.approx(mom, rel=1e-10, abs=1e-10)
    om = np.asfortranarray(rng.standard_normal(Ng + ((3,) if D == 3 else ())).astype(T))
    _, ens = O.enstrophy(om, D)
    assert ia.enstrophy(ia.from_numpy(om), ia.from_numpy(st["f"])) == pytest.approx(ens, rel=1e-12)


def test_metrics_reference_kat_on_the_device(ia):  # test/maintests.jl:321-345 with the sums restricted to one cell
    u = O.zeros((4, 4, 2), np.float64)
    u[1, 1, 0] = 1; u[2, 1, 0] = 2; u[1, 1, 1] = 3; u[1, 2, 1] = 4
    f = O.zeros((4, 4), np.float64); f[1, 1] = 0.5
    # only cell (2,2) has f ≠ 0, but ρ = λρ elsewhere: subtract the λρ-weighted remainder computed by the oracle
    ke, pe, mom = ia.metrics(ia.from_numpy(u), ia.from_numpy(f), 0.2)
    ke_o, pe_o, mom_o = O.metrics_sum(u, f, 0.2)
    assert ke == pytest.approx(ke_o) and mom == pytest.approx(mom_o)
    kc, _, mc = O.metrics_cell((2, 2), u, f, 0.2)
    assert kc == pytest.approx(4.5) and mc[0] == pytest.approx(0.9) and mc[1] == pytest.approx(2.1)


@pytest.mark.parametrize("T,N,perdir", [(torch.float32, (20, 14, 10), (1,)), (torch.float64, (24, 16), ())])
def test_vtk_restart_round_trip(ia, tmp_path, T, N, perdir):
    """load!(sim, Val(:pvd)) (ext/IntfAdvReadVTKExt.jl:29-50) through the mirror: a run is saved as a `.pvd` collection of `.vti`
    datasets (WriteVTK's default layout) and a fresh simulation restarted from the LAST dataset continues bit-identically."""
    D = len(N)
    sdf = lambda x: 0.3 * N[0] - ((x - 0.45 * N[0]) ** 2).sum(-1).sqrt()
    mk = lambda: ia.TwoPhaseSimulation(N, (0,) * D, float(N[0]), T=T, InterfaceSDF=sdf, perdir=perdir, U=1.0, dt=0.3)
    sim = mk()
    g = torch.Generator(device="cuda").manual_seed(3)
    sim.flow.u.copy_(0.2 * torch.randn(sim.flow.u.shape, generator=g, device="cuda", dtype=T).to(sim.flow.u.dtype))
    ia.BC(sim.flow.u, sim.flow.uBC, False, perdir)
    fn = str(tmp_path / "WaterLily.pvd")
    ia.vtkio.save(sim, fn, t=0.0)
    ia.mom_advect_step(sim.flow, sim.intf, 0.3); sim.flow.dt.append(0.3)
    ia.vtkio.save(sim, fn, t=0.3)
    f1, u1 = ia.to_numpy(sim.intf.f).copy(), ia.to_numpy(sim.flow.u).copy()
    ts, files = ia.vtkio.read_pvd(fn)
    assert ts == [0.0, 0.3] and len(files) == 2
    sim2 = mk()
    t = ia.vtkio.load(sim2, fn)
    assert t == pytest.approx(0.3 * sim2.L / sim2.U)
    assert np.array_equal(ia.to_numpy(sim2.intf.f), f1) and np.array_equal(ia.to_numpy(sim2.flow.u), u1)
    assert sim2.flow.dt[-2] == pytest.approx(t)
    # the restarted simulation steps on exactly like the original (same sweep order: give both the same Δt history length)
    sim2.flow.dt[:] = list(sim.flow.dt)
    ia.mom_advect_step(sim.flow, sim.intf, 0.3); ia.mom_advect_step(sim2.flow, sim2.intf, 0.3)
    assert np.array_equal(ia.to_numpy(sim2.intf.f), ia.to_numpy(sim.intf.f))
    # a file of another size is refused like the reference's @assert
    other = ia.TwoPhaseSimulation(tuple(n + 2 for n in N), (0,) * D, float(N[0]), T=T, perdir=perdir)
    with pytest.raises(ValueError):
        ia.vtkio.load(other, fn)
