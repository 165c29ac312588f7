"""Self-derived property tests for the parts of the oracle that NO reference test pins:
WH normals, SynDRoM, advectρuu1D!, limiters, MPCFL, 3-D sweeps.  They check invariants the algorithm must have;
they are not known answers from the reference.  CPU only."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests.helpers import alloc_cmom, dirO_for, inside, make_state, oracle_cmom_call, oracle_mom_advect_step


def test_quiescent_is_exact_noop():
    """maintests.jl:197-202: zero velocity => f and ρu untouched, bit for bit."""
    for N, kind, per in [((8, 8), "C1", (1, 2)), ((8, 8, 8), "C2", ())]:
        st = make_state(N, kind, np.float64, perdir=per)
        st["u"][...] = 0
        f0 = st["f"].copy(order="F")
        a = alloc_cmom(st)
        O.u2rhou(a["rhou"], st["u"], st["f"], st["lam_rho"])
        status, rep, a = oracle_cmom_call(st, st["f"], st["u"], st["u"], st["u"], a["rhou"].copy(order="F"), 0.25, dirO_for(0, st["D"]))
        assert status == 0
        assert (st["f"] == f0).all()
        assert (inside(a["rhou"], st["D"]) == 0).all()


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_tgv_droplet_mass_conservation_2d(T):
    """maintests.jl:206-215: periodic TGV droplet, Σf conserved (reference asserts rtol 1e-4 with the full solver;
    with a prescribed discretely solenoidal field the conservative sweep conserves to round-off)."""
    st = make_state((16, 16), "C4" if False else "C1", T, perdir=(1, 2))
    from interfaceadvection.jl_b200 import configs
    st["u"] = np.asfortranarray(configs.tgv((16, 16), T, U=0.25))
    O.BC(st["u"], (0, 0), False, (1, 2))
    V0 = O.sum_inside(st["f"])
    for n in range(4):
        oracle_mom_advect_step(st, st["f"], st["u"], 1.0, dirO_for(n, 2))
    assert np.isfinite(st["f"]).all()
    tol = 1e-5 if T == np.float32 else 1e-12
    assert abs(O.sum_inside(st["f"]) - V0) <= tol * V0


def test_enright_mass_conservation_3d():
    st = make_state((24, 24, 24), "C2", np.float64)
    a = alloc_cmom(st)
    V0 = O.sum_inside(st["f"])
    for n in range(10):
        s, rep = O.advectVOF(st["f"], a["ff"], a["alpha"], a["nhat"], st["u"], st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"],
                             "WH", st["perdir"], dirO_for(n, 3))
        assert s == 0, (s, rep.maxf, rep.minf)
    assert abs(O.sum_inside(st["f"]) - V0) <= 1e-12 * V0
    assert st["f"].min() >= 0 and st["f"].max() <= 1


def test_xy_symmetry_2d():
    """Transposing the problem (swap x<->y, swap velocity components, swap sweep order) transposes the result."""
    T = np.float64
    st = make_state((20, 12), "C3", T)
    f1 = st["f"].copy(order="F"); u1 = st["u"].copy(order="F")
    a = alloc_cmom(st)
    O.u2rhou(a["rhou"], u1, f1, st["lam_rho"])
    ru = a["rhou"].copy(order="F")
    oracle_cmom_call(st, f1, u1, u1, u1, ru, 1.0, (1, 2), arrays=a)
    # transposed problem
    st2 = dict(st); st2["N"] = (12, 20); st2["Ng"] = (14, 22)
    f2 = np.asfortranarray(st["f"].T.copy()); u2 = np.asfortranarray(np.stack([st["u"][..., 1].T, st["u"][..., 0].T], axis=-1))
    a2 = alloc_cmom(st2)
    O.u2rhou(a2["rhou"], u2, f2, st["lam_rho"])
    oracle_cmom_call(st2, f2, u2, u2, u2, a2["rhou"].copy(order="F"), 1.0, (2, 1), arrays=a2)
    assert np.array_equal(f2.T, f1)
    assert np.array_equal(inside(a2["rhou"][..., 1], 2).T, inside(a["rhou"][..., 0], 2))
    assert np.array_equal(inside(a2["rhou"][..., 0], 2).T, inside(a["rhou"][..., 1], 2))


def test_periodic_shift_invariance():
    """On a fully periodic box, rolling all inputs by whole cells rolls the outputs."""
    T = np.float64
    N = (16, 12, 10)
    st = make_state(N, "C2", T, perdir=(1, 2, 3))
    from interfaceadvection.jl_b200 import configs
    st["u"] = np.asfortranarray(configs.tgv(N, T, U=0.3)); O.BC(st["u"], (0, 0, 0), False, (1, 2, 3))
    f = st["f"].copy(order="F"); u = st["u"].copy(order="F")
    ru = oracle_mom_advect_step(st, f, u, 1.0, (3, 1, 2))
    sh = (5, 3, 4)

    def roll(a, vec):
        core = a[1:-1, 1:-1, 1:-1]
        core = np.roll(core, sh, axis=(0, 1, 2))
        b = a.copy(order="F"); b[1:-1, 1:-1, 1:-1] = core
        if vec:
            O.BC(b, (0, 0, 0), False, (1, 2, 3))
        else:
            O.BCf(b, (1, 2, 3))
        return b
    f2 = roll(st["f"], False); u2 = roll(st["u"], True)
    ru2 = oracle_mom_advect_step(st, f2, u2, 1.0, (3, 1, 2))
    assert np.array_equal(f2, roll(f, False))
    assert np.array_equal(inside(ru2, 3), inside(roll(ru, True), 3))


def test_limiters_are_tvd_and_consistent():
    rng = np.random.default_rng(20261017)
    for name in ["minmod", "Koren", "vanAlbada1", "Sweby", "superbee", "TVDcen", "TVDdown", "quick", "vanLeer"]:
        for _ in range(200):
            u, c, d = rng.normal(size=3)
            v = O.limiter(name, u, c, d)
            lo, hi = min(c, d), max(c, d)
            if (c - u) * (d - c) > 0:  # monotone data: face value between centre and downstream
                assert lo - 1e-14 <= v <= hi + 1e-14, name
            else:                     # extremum: upwind
                assert v == pytest.approx(c, abs=1e-14), name
        # linear data -> the second-order face value (TVDdown takes the full downstream value by definition, flow.jl:15)
        assert O.limiter(name, 1.0, 2.0, 3.0) == pytest.approx(3.0 if name == "TVDdown" else 2.5)
    assert O.limiter("upwind", 1.0, 2.0, 3.0) == 2.0
    assert O.limiter("cds", 1.0, 2.0, 4.0) == 3.0


def test_uniform_translation_keeps_velocity():
    """A uniform velocity field on a periodic box must stay uniform under CMOM: ρu_new / ρ(f_new) == U."""
    T = np.float64
    N = (24, 16)
    st = make_state(N, "C1", T, perdir=(1, 2))
    st["u"][..., 0] = 0.3; st["u"][..., 1] = -0.2
    f = st["f"]; u = st["u"]
    ru = oracle_mom_advect_step(st, f, u, 1.0, (1, 2))
    unew = O.zeros(u.shape, T)
    O.rhou2u(unew, ru, f, st["lam_rho"])
    assert np.abs(inside(unew[..., 0], 2) - 0.3).max() < 1e-12
    assert np.abs(inside(unew[..., 1], 2) + 0.2).max() < 1e-12


def test_mpcfl_closed_form():
    T = np.float64
    u = O.zeros((6, 6, 2), T); sig = O.zeros((6, 6), T)
    u[2, 2, 0] = -0.5; u[3, 2, 0] = 0.25; u[2, 3, 1] = 1.0
    # flux_out at cell (2,2): max(0,u[3,2,0]) + max(0,-u[2,2,0]) + max(0,u[2,3,1]) + 0 = 0.25+0.5+1
    dt = O.MPCFL(u, sig, nu=0.0)
    # Δt_Adv = 1/1.75 ; maxTotalFlux = max(.5,.25)+max(0,1) = 1.5 -> Δt_cVOF = 1/3
    assert dt == pytest.approx(0.8 * min(1 / 1.75, 1 / 3.0, 1.0))
