"""Pins the oracle's post-processing functions against the reference's own tests: metrics.jl known answers
(/root/reference/test/maintests.jl:321-345) and the redistancing exactness tests (:264-318).  CPU only."""
import numpy as np
import pytest

from oracle import pyoracle as O


def test_metrics_kat():  # maintests.jl:321-345
    u = O.zeros((4, 4, 2), np.float64)
    u[1, 1, 0] = 1; u[2, 1, 0] = 2; u[1, 1, 1] = 3; u[1, 2, 1] = 4
    f = O.zeros((4, 4), np.float64); f[1, 1] = 0.5
    lr, I = 0.2, (2, 2)
    ke, pe, mom = O.metrics_cell(I, u, f, lr)
    assert ke == pytest.approx(4.5) and mom[0] == pytest.approx(0.9) and mom[1] == pytest.approx(2.1)
    ke1, _, mom1 = O.metrics_cell(I, u, f, lr, U=(1.0, 1.0))
    assert ke1 == pytest.approx(2.1) and mom1[0] == pytest.approx(0.3)
    _, pe, _ = O.metrics_cell(I, u, f, lr, g=(0.0, -1.0), statWL=(0.0, 0.0))
    assert pe == pytest.approx(0.3)
    om = O.zeros((4, 4), np.float64); om[1, 1] = 1; om[2, 1] = 2; om[1, 2] = 3; om[2, 2] = 4
    assert O.enstrophy(om, 2, I)[0] == pytest.approx(3.75)
    om3 = O.zeros((4, 4, 4, 3), np.float64)
    om3[1, 1, 1, 0] = 1; om3[1, 2, 1, 0] = 2; om3[1, 1, 2, 0] = 3; om3[1, 2, 2, 0] = 4
    om3[1, 1, 1, 1] = 1; om3[1, 1, 2, 1] = 2; om3[2, 1, 1, 1] = 3; om3[2, 1, 2, 1] = 4
    om3[1, 1, 1, 2] = 1; om3[2, 1, 1, 2] = 2; om3[1, 2, 1, 2] = 3; om3[2, 2, 1, 2] = 4
    assert O.enstrophy(om3, 3, (2, 2, 2))[0] == pytest.approx(11.25)


def test_metrics_sums_are_the_sums_of_the_cells():
    rng = np.random.default_rng(4)
    Ng = (7, 6, 5)
    u = np.asfortranarray(rng.standard_normal(Ng + (3,))); f = np.asfortranarray(rng.uniform(0, 1, Ng))
    ke, pe, mom = O.metrics_sum(u, f, 0.01, U=(0.1, 0.2, 0.3), g=(0, -9.81, 0), statWL=(0, 2.0, 0))
    ke2 = pe2 = 0.0; mom2 = [0.0] * 3
    for i in range(2, Ng[0]):
        for j in range(2, Ng[1]):
            for k in range(2, Ng[2]):
                a, b, m = O.metrics_cell((i, j, k), u, f, 0.01, U=(0.1, 0.2, 0.3), g=(0, -9.81, 0), statWL=(0, 2.0, 0))
                ke2 += a; pe2 += b; mom2 = [x + y for x, y in zip(mom2, m)]
    assert ke == pytest.approx(ke2, rel=1e-13) and pe == pytest.approx(pe2, rel=1e-13)
    assert mom == pytest.approx(mom2, rel=1e-12)


def test_compute_l_vanishes_on_a_distance_ramp():  # maintests.jl:267-277
    N = (18, 18)
    phi = np.asfortranarray(np.array([[i - 2.5 - 8 for _ in range(1, N[1] + 1)] for i in range(1, N[0] + 1)], dtype=np.float64))
    pini = phi.copy(order="F"); L = O.zeros(N, np.float64)
    O.computeL(L, phi, pini, ())
    assert np.abs(L[1:-1, 1:-1]).max() < 1e-12
    O.computeL(L, phi, pini, (2,))
    assert np.abs(L[1:-1, 1:-1]).max() < 1e-12


def test_stage_sequence_on_uniform_l():  # maintests.jl:279-292
    phid = np.asfortranarray(np.array([[i + j for j in range(1, 7)] for i in range(1, 7)], dtype=np.float64))
    dtau, Lval = 0.01, 1 - np.sqrt(2)
    Lbuf = O.zeros((6, 6), np.float64); phi0 = phid.copy(order="F")
    # the reference passes ϕd itself as ϕini (aliased); the oracle reads ϕini while it writes ϕ, exactly like the reference's loops
    for al in (0.0, 0.75, 1 / 3):
        O.redistaningStage(phid, phi0, phid, Lbuf, dtau, al)
    assert np.abs(phid[1:-1, 1:-1] - (phi0[1:-1, 1:-1] + dtau * Lval)).max() < 0.05


def test_redistancing_of_a_planar_interface():  # maintests.jl:304-318
    N = (16, 16)
    Ng = (18, 18)
    f = O.zeros(Ng, np.float64); al = O.zeros(Ng, np.float64); nh = O.zeros(Ng + (2,), np.float64)
    O.applyVOF(f, al, nh, lambda x: x[..., 0] - 8)
    O.BCf(f, (2,))
    phi = np.asfortranarray(2 * f - 1); pini = phi.copy(order="F"); phi0 = O.zeros(Ng, np.float64); L = O.zeros(Ng, np.float64)
    O.redistaning(phi, phi0, pini, L, d=4, dtau=0.05, perdir=(2,))
    assert np.isfinite(phi).all()
    col = phi[1:-1, 8]
    assert np.all(np.diff(col) <= 0)
    assert np.array_equal(phi[:, 0], phi[:, -2]) and np.array_equal(phi[:, -1], phi[:, 1])
    for ix in range(8, 12):
        assert phi[ix - 1, 8] == pytest.approx(-(ix - 1.5 - 8), abs=0.01)


def test_gradphi2_upwinding():
    """Self-derived: on a linear profile of slope m the one-sided estimate is m² when the upwind side exists, else 0."""
    assert O.gradphi2(-2, -1, 0, 1, 2, 1.0) == pytest.approx(1.0)      # ϕ rising, s>0: information comes from the left
    assert O.gradphi2(-2, -1, 0, 1, 2, -1.0) == pytest.approx(1.0)
    assert O.gradphi2(2, 1, 0, 1, 2, 1.0) == 0.0                        # local minimum, s>0: no upwind side
    assert O.gradphi2(-2, -1, 0, -1, -2, 1.0) == 0.0                    # symmetric local maximum: wᴿ + wᴸ = 0, neither branch
    assert O.gradphi2(-4, -2, 0, -1, -2, 1.0) == pytest.approx(4.0)  # asymmetric maximum: the left one-sided slope 2 (its curvature correction is minmod(-3,0) = 0)
