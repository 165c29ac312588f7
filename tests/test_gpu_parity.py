"""Parity of the sm_100a path (through the C ABI of libifadv_b200.so) against the CPU oracle on the same inputs.
Tolerances are the ones BASELINE.json's north_star states: max|Δf|, |Δρu| <= 1e-12 (Float64), <= 1e-5 (Float32)
after one step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as O  # noqa: E402
from tests.helpers import alloc_cmom, dirO_for, inside, make_state, oracle_cmom_call, oracle_mom_advect_step  # noqa: E402

TOL = {np.float32: 1e-5, np.float64: 1e-12}


@pytest.fixture(scope="module")
def ia():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import interfaceadvection.jl_b200 as ia

    return ia


def dev_arrays(ia, st, a):
    d = {k: ia.from_numpy(v) for k, v in a.items()}
    return d


def run_cuda_cmom(ia, st, f, u1, u2, uOld, rhou, dt, dirO, lam="Koren", scheme="WH"):
    a = alloc_cmom(st)
    a["rhou"][...] = rhou
    d = dev_arrays(ia, st, a)
    fd, u1d, u2d, uod = ia.from_numpy(f), ia.from_numpy(u1), ia.from_numpy(u2), ia.from_numpy(uOld)
    status = ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], u1d, u2d, dt, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"],
                               d["nhat"], uod, d["alpha"], d["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO)
    return status, ia.to_numpy(fd), ia.to_numpy(d["rhou"])


CMOM_CASES = [
    # N, kind, perdir, uBC
    ((24, 16), "C1", (), (0, 0)),
    ((24, 16), "C3", (1, 2), (0, 0)),
    ((20, 12), "C3", (2,), (0, 0)),
    ((16, 12, 10), "C2", (), (0, 0, 0)),
    ((16, 12, 10), "C3", (1, 2, 3), (0, 0, 0)),
    ((12, 16, 10), "C4", (1, 2), (0, 0, 0)),
    ((40, 9, 7), "C3", (3,), (0, 0, 0)),
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir,uBC", CMOM_CASES)
def test_cmom_sweeps_match_oracle(ia, T, N, kind, perdir, uBC):
    D = len(N)
    for dirO in ([(1, 2), (2, 1)] if D == 2 else [(3, 1, 2), (1, 2, 3), (2, 3, 1)]):
        st = make_state(N, kind, T, perdir=perdir, uBC=uBC)
        a0 = alloc_cmom(st)
        O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
        rhou0 = a0["rhou"].copy(order="F")
        f_o = st["f"].copy(order="F")
        so, rep, ao = oracle_cmom_call(st, f_o, st["u"], st["u"], st["u"], rhou0, 1.0, dirO)
        sc, f_c, ru_c = run_cuda_cmom(ia, st, st["f"], st["u"], st["u"], st["u"], rhou0, 1.0, dirO)
        assert sc == so
        assert np.abs(f_c - f_o).max() <= TOL[T], (dirO, np.abs(f_c - f_o).max())
        scale = max(1.0, np.abs(inside(ao["rhou"], D)).max())
        err = np.abs(inside(ru_c, D) - inside(ao["rhou"], D)).max()
        assert err <= TOL[T] * scale, (dirO, err)
        # ghosts of f are refreshed exactly like BCf!
        g = f_c.copy(order="F"); O.BCf(g, st["perdir"])
        assert np.array_equal(g, f_c)


@pytest.mark.parametrize("lam", ["upwind", "minmod", "Koren", "vanAlbada1", "Sweby", "superbee", "TVDcen", "TVDdown", "quick", "vanLeer", "cds"])
def test_cmom_all_limiters(ia, lam):
    T = np.float64
    st = make_state((16, 12, 10), "C2", T, perdir=(1,))
    a0 = alloc_cmom(st)
    O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
    rhou0 = a0["rhou"].copy(order="F")
    f_o = st["f"].copy(order="F")
    so, rep, ao = oracle_cmom_call(st, f_o, st["u"], st["u"], st["u"], rhou0, 1.0, (3, 1, 2), lam=lam)
    sc, f_c, ru_c = run_cuda_cmom(ia, st, st["f"], st["u"], st["u"], st["u"], rhou0, 1.0, (3, 1, 2), lam=lam)
    assert np.abs(f_c - f_o).max() <= 1e-12
    assert np.abs(inside(ru_c, 3) - inside(ao["rhou"], 3)).max() <= 1e-12


def test_cmom_inflow_dirichlet(ia):
    """uBC ≠ 0 with walls: exercises the Dirichlet planes of BC! on u★ and ρuf and the ϕuL/ϕuR boundary stencils.
    The interface stays away from the inflow plane (stale-ghost quirk of the reference)."""
    for T in (np.float64, np.float32):
        st = make_state((24, 16), "C1", T, perdir=(2,), uBC=(0.2, 0.0), scale_u=0.5)
        a0 = alloc_cmom(st)
        O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
        rhou0 = a0["rhou"].copy(order="F")
        f_o = st["f"].copy(order="F")
        so, rep, ao = oracle_cmom_call(st, f_o, st["u"], st["u"], st["u"], rhou0, 1.0, (1, 2))
        sc, f_c, ru_c = run_cuda_cmom(ia, st, st["f"], st["u"], st["u"], st["u"], rhou0, 1.0, (1, 2))
        assert np.abs(f_c - f_o).max() <= TOL[T]
        assert np.abs(inside(ru_c, 2) - inside(ao["rhou"], 2)).max() <= TOL[T]


SCHEMES = ["WH", "WY", "Column", "PCD", "SLIC", "MYC", "Y", "CD", "XYLIC"]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_pure_vof_all_normal_schemes(ia, T, scheme):
    for N, kind, perdir in [((32, 32), "C1", ()), ((14, 12, 10), "C2", (1, 2, 3))]:
        D = len(N)
        st = make_state(N, kind, T, perdir=perdir)
        a = alloc_cmom(st)
        f_o = st["f"].copy(order="F")
        dirO = dirO_for(0, D)
        so, rep = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], st["u"], st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"], scheme,
                              perdir, dirO)
        d = dev_arrays(ia, st, alloc_cmom(st))
        fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
        sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], scheme, perdir, dirO)
        assert sc == so
        assert np.abs(ia.to_numpy(fd) - f_o).max() <= TOL[T], scheme
        # ρuf (mass flux·δt) is part of advect!'s visible state on inside_uWB faces
        ruf_c, ruf_o = ia.to_numpy(d["rhouf"]), a["rhouf"]
        for j in range(D):
            sl = [slice(1, -1)] * D; sl[j] = slice(1, None)
            assert np.abs(ruf_c[tuple(sl) + (j,)] - ruf_o[tuple(sl) + (j,)]).max() <= TOL[T], (scheme, j)
        assert np.array_equal(ia.to_numpy(d["cbar"])[tuple([slice(1, -1)] * D)], a["cbar"][tuple([slice(1, -1)] * D)])


def test_zalesak_128_config1_many_steps(ia):
    """BASELINE config 1 at full size (2-D Zalesak 128², Float64, pure VOF): 64 steps against the oracle."""
    T = np.float64
    st = make_state((128, 128), "C1", T)
    a = alloc_cmom(st)
    f_o = st["f"].copy(order="F")
    d = dev_arrays(ia, st, alloc_cmom(st))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    V0 = O.sum_inside(f_o)
    worst = 0.0
    for n in range(64):
        dirO = dirO_for(n, 2)
        O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], st["u"], st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", (), dirO)
        ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), dirO)
        if n == 0:
            assert np.abs(ia.to_numpy(fd) - f_o).max() <= 1e-12
        worst = max(worst, np.abs(ia.to_numpy(fd) - f_o).max())
    assert worst <= 1e-9, worst  # 64 steps of accumulated libm-vs-libdevice ulps
    assert abs(ia.sum_inside(fd) - V0) <= 1e-11 * V0


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_mom_advect_step_and_host_entry(ia, T):
    """Transport part of MPFMomStep! (flow.jl:61,69-70,74,89-92) through the Python mirror and through the
    host-buffer C entry point, both against the oracle."""
    from interfaceadvection.jl_b200 import _lib
    st = make_state((16, 12, 10), "C3", T, perdir=(3,))
    f_o = st["f"].copy(order="F")
    ru_o = oracle_mom_advect_step(st, f_o, st["u"], 1.0, (3, 1, 2))
    # host-buffer entry point
    f_h = st["f"].copy(order="F"); ru_h = np.zeros_like(st["u"], order="F")
    ctx = ia.Context(st["Ng"], np.dtype(T).name, 0)
    rep = ia.Report()
    rc = ctx.mom_advect_step_host(f_h.ctypes.data, st["u"].ctypes.data, ru_h.ctypes.data, 1.0, st["lam_rho"], ia.LIMITERS["Koren"],
                                  ia.NORMAL_SCHEMES["WH"], st["uBC"], st["perdir"], (3, 1, 2), rep)
    assert rc == 0
    assert np.abs(f_h - f_o).max() <= TOL[T]
    assert np.abs(inside(ru_h, 3) - inside(ru_o, 3)).max() <= TOL[T]
    assert ctx.launches > 0


def test_utilities_match_oracle(ia):
    rng = np.random.default_rng(20261017)
    for T in (np.float32, np.float64):
        for Ng, per in [((9, 7), (1,)), ((8, 6, 7), (2, 3)), ((8, 6, 7), ())]:
            D = len(Ng)
            f = np.asfortranarray(rng.random(Ng).astype(T)); u = np.asfortranarray(rng.normal(size=Ng + (D,)).astype(T))
            # BCf!
            g_o = f.copy(order="F"); O.BCf(g_o, per)
            g_d = ia.from_numpy(f); ia.BCf(g_d, per)
            assert np.array_equal(ia.to_numpy(g_d), g_o)
            # BC!
            A = (0.3, -0.2, 0.1)[:D]
            for saveexit in (False, True):
                b_o = u.copy(order="F"); O.BC(b_o, A, saveexit, per)
                b_d = ia.from_numpy(u); ia.BC(b_d, A, saveexit, per)
                assert np.array_equal(ia.to_numpy(b_d), b_o), (Ng, per, saveexit)
            # u2ρu! / ρu2u!
            r_o = O.zeros(Ng + (D,), T); O.u2rhou(r_o, u, g_o, 0.1)
            r_d = ia.jl_zeros(Ng + (D,), getattr(torch, np.dtype(T).name)); ia.u2rhou(r_d, ia.from_numpy(u), ia.from_numpy(g_o), 0.1)
            assert np.array_equal(inside(ia.to_numpy(r_d), D), inside(r_o, D))
            v_o = O.zeros(Ng + (D,), T); O.rhou2u(v_o, r_o, g_o, 0.1)
            v_d = ia.jl_zeros(Ng + (D,), getattr(torch, np.dtype(T).name)); ia.rhou2u(v_d, ia.from_numpy(r_o), ia.from_numpy(g_o), 0.1)
            assert np.array_equal(inside(ia.to_numpy(v_d), D), inside(v_o, D))
            # MPCFL, sum
            sig = O.zeros(Ng, T)
            dt_o = O.MPCFL(u, sig, nu=0.01, mu=0.01, eta=0.5)
            dt_d = ia.context_for(g_d).mpcfl(0, ia.from_numpy(u).data_ptr(), nu=0.01, mu=0.01, eta=0.5)
            assert dt_d == pytest.approx(dt_o, rel=1e-6 if T == np.float32 else 1e-13)
            assert ia.sum_inside(g_d) == pytest.approx(O.sum_inside(g_o), rel=1e-12)


def test_applyvof_matches_oracle_and_reference_kat(ia):
    """applyVOF! on the device: the reference's own fRef (maintests.jl:131-136) and the oracle on a sphere."""
    f = ia.jl_zeros((4, 4), torch.float64); al = ia.jl_zeros((4, 4), torch.float64); nh = ia.jl_zeros((4, 4, 2), torch.float64)
    ia.applyVOF(f, al, nh, lambda x: (-x[..., 0] - 3 * x[..., 1] + 4.5) / 10 ** 0.5)
    fRef = np.array([[0, 0, 0, 0], [0, 0, 2 / 3, 0], [0, 1 / 24, 23 / 24, 0], [0, 0, 0, 0]])
    assert np.allclose(ia.to_numpy(f), fRef, rtol=1.5e-8, atol=1e-9)
    from interfaceadvection.jl_b200 import configs
    sdf = configs.sdf_sphere([7.3, 6.1, 5.2], 3.7)
    Ng = (14, 12, 11)
    fo = O.zeros(Ng, np.float64); ao = O.zeros(Ng, np.float64); no = O.zeros(Ng + (3,), np.float64)
    O.applyVOF(fo, ao, no, sdf)
    fd = ia.jl_zeros(Ng, torch.float64); ad = ia.jl_zeros(Ng, torch.float64); nd = ia.jl_zeros(Ng + (3,), torch.float64)
    ia.applyVOF(fd, ad, nd, sdf)
    assert np.abs(ia.to_numpy(fd) - fo).max() <= 1e-12


def test_simulation_surface_and_mass_conservation(ia):
    """TwoPhaseSimulation / sim_step! surface (InterfaceAdvection.jl:64-107): periodic TGV droplet like
    test/helper.jl:7-13, prescribed velocity, 20 steps; Σf conserved to round-off, dirO rotates with length(Δt)."""
    from interfaceadvection.jl_b200 import configs
    N = (32, 32, 32)
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 32.0, T=torch.float64, lam_rho=0.1, InterfaceSDF=configs.sdf_sphere([16, 16, 16], 8.0),
                                perdir=(1, 2, 3), U=1.0, dt=1.0)
    sim.flow.u.copy_(ia.from_numpy(configs.tgv(N, np.float64, U=0.25)))
    ia.BC(sim.flow.u, (0, 0, 0), False, (1, 2, 3))
    V0 = ia.sum_inside(sim.intf.f)
    n0 = len(sim.flow.dt)
    for _ in range(20):
        ia.sim_step(sim)
    assert len(sim.flow.dt) == n0 + 20
    assert torch.isfinite(sim.intf.f).all() and torch.isfinite(sim.intf.rhou).all()
    assert abs(ia.sum_inside(sim.intf.f) - V0) <= 1e-12 * V0


def test_nan_is_fatal(ia):
    """reportFillError: NaN -> error("NaN!") (advection.jl:148) -> status -1 -> IfadvError."""
    st = make_state((12, 12), "C1", np.float64)
    st["f"][5, 5] = np.nan
    d = dev_arrays(ia, st, alloc_cmom(st))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    with pytest.raises(ia.IfadvError, match="NaN"):
        ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2))


def test_rejects_unsupported(ia):
    st = make_state((12, 12), "C1", np.float64)
    d = dev_arrays(ia, st, alloc_cmom(st))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    with pytest.raises(ia.IfadvError):
        ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"], d["nhat"],
                          ud, d["alpha"], d["drho"], 1e-3, "Koren", "WH", (0, 0), (), True, (1, 2))
    with pytest.raises(ia.IfadvError):
        ia.BC(ud, lambda i, x, t: 0.0)
    with pytest.raises(ia.IfadvError):
        ia.BCf(torch.zeros(4, 4, dtype=torch.float64))  # CPU tensor: no fallback


def _fresh_context_env(ia, kernel):
    """Contexts read IFADV_KERNEL at creation: drop the cache so the next call builds one with the requested kernel family."""
    import os
    from interfaceadvection.jl_b200 import api
    api._contexts.clear()
    if kernel is None:
        os.environ.pop("IFADV_KERNEL", None)
    else:
        os.environ["IFADV_KERNEL"] = kernel


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_kernel_families_agree(ia, T):
    """The generations of the fused sweep (v1 tile, plane-marching, register-marching, lean register-marching) implement the same
    arithmetic: bit-identical in Float64, and within a few ulp in Float32 (the v1 tile kernel is built exact, the
    marching kernels with IFADV_FAST_F32)."""
    st = make_state((70, 40, 36), "C3", T, perdir=(2,), uBC=(0.0, 0.0, 0.0))
    a0 = alloc_cmom(st)
    O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
    rhou0 = a0["rhou"].copy(order="F")
    res = {}
    try:
        for kern in ("tile", "march", None):
            _fresh_context_env(ia, kern)
            sc, f_c, ru_c = run_cuda_cmom(ia, st, st["f"], st["u"], st["u"], st["u"], rhou0, 1.0, (3, 1, 2))
            res[kern] = (f_c, ru_c)
    finally:
        _fresh_context_env(ia, None)
    f_o = st["f"].copy(order="F")
    so, rep, ao = oracle_cmom_call(st, f_o, st["u"], st["u"], st["u"], rhou0, 1.0, (3, 1, 2))
    for kern, (f_c, ru_c) in res.items():
        assert np.abs(f_c - f_o).max() <= TOL[T], kern
        assert np.abs(inside(ru_c, 3) - inside(ao["rhou"], 3)).max() <= TOL[T], kern
    if T == np.float64:
        assert np.array_equal(res["tile"][0], res[None][0]) and np.array_equal(res["march"][0], res[None][0])
        assert np.array_equal(inside(res["tile"][1], 3), inside(res[None][1], 3))
        assert np.array_equal(inside(res["march"][1], 3), inside(res[None][1], 3))


def test_full_size_properties_enright_256(ia):
    """BASELINE config 2 at full size (3-D Enright/LeVeque 256³, Float32, pure VOF, walls): size-independent properties.
    The velocity is the discrete curl of a vector potential, so Σf must be conserved to round-off, f must stay in
    [0,1] and the ghost layer must equal BCf! of the interior."""
    from interfaceadvection.jl_b200 import configs
    N = (256, 256, 256)
    dev = torch.device("cuda", 0)
    case = configs.make_case("C2_enright_256", device=dev)
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 256.0, T=torch.float32, InterfaceSDF=case["sdf"], perdir=(), U=1.0, dt=1.0, device=dev)
    sim.flow.u.copy_(case["u"]); ia.BC(sim.flow.u, (0, 0, 0), False, ()); sim.flow.u0.copy_(sim.flow.u)
    V0 = ia.sum_inside(sim.intf.f)
    for n in range(6):
        st = ia.advect(sim.flow, sim.intf)
        sim.flow.dt.append(1.0)
        assert st == 0
    V1 = ia.sum_inside(sim.intf.f)
    assert abs(V1 - V0) <= 2e-6 * V0, (V0, V1)          # Float32 round-off over 18 sweeps of 16.7 M cells
    f = sim.intf.f
    assert float(f.min()) >= 0.0 and float(f.max()) <= 1.0
    g = f.clone(memory_format=torch.preserve_format); ia.BCf(g, ())
    assert torch.equal(g, f)


def test_full_size_properties_dambreak(ia):
    """BASELINE config 3 (3-D dam break 512x256x256, Float32, walls, λρ = 1e-3, full CMOM + SynDRoM): mass conservation,
    boundedness and finite momentum over two CMOM advection steps at full size."""
    from interfaceadvection.jl_b200 import configs
    N = (512, 256, 256)
    dev = torch.device("cuda", 0)
    case = configs.make_case("C3_dambreak_512x256x256", device=dev)
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 512.0, T=torch.float32, lam_rho=1e-3, InterfaceSDF=case["sdf"], perdir=(), U=1.0, dt=1.0,
                                device=dev)
    sim.flow.u.copy_(case["u"]); ia.BC(sim.flow.u, (0, 0, 0), False, ())
    V0 = ia.sum_inside(sim.intf.f)
    for _ in range(2):
        ia.mom_advect_step(sim.flow, sim.intf, 1.0, check=True)
        sim.flow.dt.append(1.0)
    V1 = ia.sum_inside(sim.intf.f)
    assert abs(V1 - V0) <= 2e-6 * V0
    assert float(sim.intf.f.min()) >= 0.0 and float(sim.intf.f.max()) <= 1.0
    assert bool(torch.isfinite(sim.intf.rhou[1:-1, 1:-1, 1:-1]).all())
    del sim
    torch.cuda.empty_cache()


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir,uBC", [((40, 24, 20), "C3", (), (0, 0, 0)), ((32, 24, 20), "C4", (1, 2, 3), (0, 0, 0)),
                                               ((36, 20, 18), "C2", (2,), (0.1, 0.0, 0.0)), ((24, 16), "C1", (), (0, 0))])
def test_fused_entry_equals_separate_calls(ia, T, N, kind, perdir, uBC):
    """ifadv_u2rhou_advect_vof_rhouu == f.=f_src; u2ρu!; BC!; advectVOFρuu!  (bit-identical in Float64, <= tol in Float32),
    and both match the oracle's transport half of MPFMomStep!."""
    D = len(N)
    TT = getattr(torch, np.dtype(T).name)
    st = make_state(N, kind, T, perdir=perdir, uBC=uBC, scale_u=0.6)
    f_o = st["f"].copy(order="F")
    dirO = dirO_for(0, D)
    ru_o = oracle_mom_advect_step(st, f_o, st["u"], 1.0, dirO)
    out = {}
    for fused in (False, True):
        flow = ia.Flow(st["N"], st["uBC"], T=TT, dt=1.0, perdir=st["perdir"])
        intf = ia.cVOF(st["N"], T=TT, lam_rho=st["lam_rho"], perdir=st["perdir"])
        flow.u.copy_(ia.from_numpy(st["u"])); intf.f.copy_(ia.from_numpy(st["f"]))
        ia.mom_advect_step(flow, intf, 1.0, fused=fused)
        out[fused] = (ia.to_numpy(intf.f), ia.to_numpy(intf.rhou), ia.to_numpy(intf.f0))
    for fused in (False, True):
        assert np.abs(out[fused][0] - f_o).max() <= TOL[T]
        assert np.abs(inside(out[fused][1], D) - inside(ru_o, D)).max() <= TOL[T] * max(1.0, np.abs(ru_o).max())
    if T == np.float64:
        assert np.array_equal(out[True][0], out[False][0])
        assert np.array_equal(inside(out[True][1], D), inside(out[False][1], D))
        assert np.array_equal(out[True][2], out[False][2])


@pytest.mark.parametrize("N,perdir", [((6, 5, 4), (1, 2, 3)), ((5, 4, 6), ()), ((33, 3, 3), (1,)), ((4, 4), (1, 2)), ((3, 7), ())])
def test_tiny_and_ragged_grids(ia, N, perdir):
    """Edge cases: grids smaller than a tile / a chunk warm-up / a periodic stencil reach, odd extents."""
    T = np.float64
    D = len(N)
    rng = np.random.default_rng(20261017)
    Ng = tuple(n + 2 for n in N)
    f = np.asfortranarray(np.clip(rng.random(Ng) * 1.6 - 0.3, 0, 1)); O.BCf(f, perdir)
    u = np.asfortranarray(rng.normal(size=Ng + (D,)) * 0.08); O.BC(u, (0,) * D, False, perdir)
    st = dict(N=N, D=D, Ng=Ng, dtype=T, perdir=tuple(perdir), uBC=(0.0,) * D, f=f, u=u, lam_rho=0.01)
    f_o = f.copy(order="F")
    dirO = dirO_for(1, D)
    ru_o = oracle_mom_advect_step(st, f_o, u, 1.0, dirO)
    TT = torch.float64
    flow = ia.Flow(N, st["uBC"], T=TT, dt=1.0, perdir=perdir)
    flow.dt.append(1.0)
    intf = ia.cVOF(N, T=TT, lam_rho=0.01, perdir=perdir)
    flow.u.copy_(ia.from_numpy(u)); intf.f.copy_(ia.from_numpy(f))
    ia.mom_advect_step(flow, intf, 1.0)
    assert np.abs(ia.to_numpy(intf.f) - f_o).max() <= 1e-12
    assert np.abs(inside(ia.to_numpy(intf.rhou), D) - inside(ru_o, D)).max() <= 1e-12


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,chunk", [((16, 12, 40), 8), ((12, 10, 37), 16), ((8, 8, 20), 3)])
def test_host_entry_pipelined(ia, T, N, chunk, monkeypatch):
    """The z-slab pipeline of ifadv_mom_advect_step_host (H2D / step / D2H of consecutive slabs overlap; 8 overlap planes per
    artificial slab end) returns exactly what the single-pass entry returns, and both match the oracle."""
    st = make_state(N, "C3", T, perdir=(1,))
    f_o = st["f"].copy(order="F")
    ru_o = oracle_mom_advect_step(st, f_o, st["u"], 1.0, (2, 3, 1))
    TT = getattr(torch, np.dtype(T).name)

    def pinned_like(a):
        t = torch.empty(a.size, dtype=TT, pin_memory=True)
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return t, v

    out = {}
    for mode, env in (("single", "0"), ("pipelined", str(chunk))):
        monkeypatch.setenv("IFADV_HOST_CHUNK", env)
        ft, fv = pinned_like(st["f"]); ut, uv = pinned_like(st["u"]); rt, rv = pinned_like(np.zeros_like(st["u"]))
        ctx = ia.Context(st["Ng"], np.dtype(T).name, 0)
        rep = ia.Report()
        rc = ctx.mom_advect_step_host(ft.data_ptr(), ut.data_ptr(), rt.data_ptr(), 1.0, st["lam_rho"], ia.LIMITERS["Koren"],
                                      ia.NORMAL_SCHEMES["WH"], st["uBC"], st["perdir"], (2, 3, 1), rep)
        assert rc == 0, mode
        assert ctx.launches > 0
        out[mode] = (fv.copy(), inside(rv, 3).copy(), ctx.launches)
    assert out["pipelined"][2] > out["single"][2]  # several slabs were launched
    assert np.array_equal(out["single"][0], out["pipelined"][0])
    assert np.array_equal(out["single"][1], out["pipelined"][1])
    assert np.abs(out["pipelined"][0] - f_o).max() <= TOL[T]
    assert np.abs(out["pipelined"][1] - inside(ru_o, 3)).max() <= TOL[T]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("Ng", [(7, 5, 3), (18, 14, 12), (9, 7)])
def test_axpby_midpoint(ia, T, Ng):
    """f⁰ = (f⁰+f)/2 (flow.jl:74) and the general a*x+b*y form, vector body + scalar tail, in place and out of place."""
    rng = np.random.default_rng(20261017)
    x = np.asfortranarray(rng.random(Ng).astype(T)); y = np.asfortranarray(rng.random(Ng).astype(T))
    ctx = ia.context_for(ia.from_numpy(x))
    xd, yd = ia.from_numpy(x), ia.from_numpy(y)
    od = ia.from_numpy(np.zeros_like(x))
    s = torch.cuda.current_stream().cuda_stream
    ctx.axpby(s, od.data_ptr(), 0.5, xd.data_ptr(), 0.5, yd.data_ptr())
    assert np.array_equal(ia.to_numpy(od), (x + y) * T(0.5))
    ctx.axpby(s, od.data_ptr(), 0.25, xd.data_ptr(), 2.0, yd.data_ptr())
    assert np.allclose(ia.to_numpy(od), T(0.25) * x + T(2.0) * y, rtol=4 * np.finfo(T).eps, atol=0)
    ctx.axpby(s, xd.data_ptr(), 0.5, xd.data_ptr(), 0.5, yd.data_ptr())  # in place, as MPFMomStep! uses it
    assert np.array_equal(ia.to_numpy(xd), (x + y) * T(0.5))


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_overlapped_step_equals_sequential(ia, T):
    """mom_advect_step with the field copies / midpoint on a second stream (overlap=True, the default without a projection hook)
    returns bit for bit what the strictly sequential order of MPFMomStep! returns -- f, f⁰, u⁰ and ρu -- over several steps."""
    st = make_state((40, 24, 20), "C3", T, perdir=(2,))
    TT = getattr(torch, np.dtype(T).name)
    res = {}
    for ov in (False, True):
        flow = ia.Flow(st["N"], st["uBC"], T=TT, dt=1.0, perdir=st["perdir"])
        intf = ia.cVOF(st["N"], T=TT, lam_rho=st["lam_rho"], perdir=st["perdir"])
        flow.u.copy_(ia.from_numpy(st["u"])); intf.f.copy_(ia.from_numpy(st["f"]))
        flow.u0.zero_()  # stale on purpose: the step has to refresh it (flow.jl:61)
        for n in range(3):
            ia.mom_advect_step(flow, intf, 1.0, overlap=ov)
            flow.dt.append(1.0)
            flow.u.mul_(0.9)  # a changing velocity between steps, as a projection would leave it
        torch.cuda.synchronize()
        res[ov] = [ia.to_numpy(x) for x in (intf.f, intf.f0, flow.u0, intf.rhou)]
    for x, y in zip(res[False], res[True]):
        assert np.array_equal(x, y)


def test_full_size_properties_bubble_512(ia):
    """BASELINE config 4 at full size (rising-bubble grid 512³, Float32, periodic x/y, λρ = 1e-3, full CMOM + SynDRoM; the bench
    workload): mass conservation, boundedness, finite momentum, ghosts == BCf!(interior), and the shift invariance the periodic
    directions offer -- the state shifted by (37, 11) cells in x, y gives the shifted result bit for bit."""
    from interfaceadvection.jl_b200 import configs
    N = (512, 512, 512)
    dev = torch.device("cuda", 0)
    case = configs.make_case("C4_bubble_512", device=dev)
    per = (1, 2)

    def run(f0, u0, steps):
        sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 512.0, T=torch.float32, lam_rho=1e-3, perdir=per, U=1.0, dt=1.0, device=dev)
        sim.intf.f.copy_(f0); ia.BCf(sim.intf.f, per)
        sim.flow.u.copy_(u0); ia.BC(sim.flow.u, (0, 0, 0), False, per)
        V0 = ia.sum_inside(sim.intf.f)
        for _ in range(steps):
            ia.mom_advect_step(sim.flow, sim.intf, 1.0, check=True)
            sim.flow.dt.append(1.0)
        V1 = ia.sum_inside(sim.intf.f)
        f, ru = sim.intf.f.clone(memory_format=torch.preserve_format), sim.intf.rhou[1:-1, 1:-1, 1:-1].clone()
        del sim
        torch.cuda.empty_cache()
        return V0, V1, f, ru

    base = ia.TwoPhaseSimulation(N, (0, 0, 0), 512.0, T=torch.float32, lam_rho=1e-3, InterfaceSDF=case["sdf"], perdir=per, U=1.0, dt=1.0,
                                 device=dev)
    f0 = base.intf.f.clone(memory_format=torch.preserve_format)
    u0 = case["u"].clone(memory_format=torch.preserve_format)
    del base, case
    torch.cuda.empty_cache()
    V0, V1, f, ru = run(f0, u0, 2)
    assert abs(V1 - V0) <= 2e-6 * V0
    assert float(f.min()) >= 0.0 and float(f.max()) <= 1.0
    assert bool(torch.isfinite(ru).all())
    g = f.clone(memory_format=torch.preserve_format); ia.BCf(g, per)
    assert torch.equal(g, f)
    # periodic shift invariance (interior cells; the ghost layers are rebuilt by BCf! / BC!)
    sx, sy = 37, 11  # not a multiple of the 32 x 16 tile: every cell meets different tile / warp / halo roles
    fs = f0.clone(memory_format=torch.preserve_format); us = u0.clone(memory_format=torch.preserve_format)
    fs[1:-1, 1:-1, :] = torch.roll(f0[1:-1, 1:-1, :], (sx, sy), (0, 1))
    us[1:-1, 1:-1, :, :] = torch.roll(u0[1:-1, 1:-1, :, :], (sx, sy), (0, 1))
    _, _, f2, ru2 = run(fs, us, 2)
    assert torch.equal(f2[1:-1, 1:-1, 1:-1], torch.roll(f[1:-1, 1:-1, 1:-1], (sx, sy), (0, 1)))
    assert torch.equal(ru2, torch.roll(ru, (sx, sy), (0, 1)))


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,perdir", [((40, 24, 20), (2,)), ((33, 17, 19), (1, 2, 3))])
def test_single_velocity_array_kernels_equal_two_array_kernels(ia, T, N, perdir):
    """advectVOFρuu! called with ONE array for u¹ and u² (what MPFMomStep! does, flow.jl:92) runs the instantiations without the
    second velocity stream; handing the same values as two distinct arrays runs the general ones.  Bit-identical f, ρu, c̄."""
    st = make_state(N, "C3", T, perdir=perdir)
    rng = np.random.default_rng(20261017)
    uOld = np.asfortranarray(st["u"] * T(0.9) + T(0.01) * rng.standard_normal(st["u"].shape).astype(T))
    O.BC(uOld, st["uBC"], False, perdir)
    rhou = O.zeros(st["u"].shape, T); O.u2rhou(rhou, uOld, st["f"], st["lam_rho"]); O.BC(rhou, st["uBC"], False, perdir)
    outs = []
    for alias in (True, False):
        a = alloc_cmom(st); a["rhou"][...] = rhou
        d = dev_arrays(ia, st, a)
        fd, ud, uod = ia.from_numpy(st["f"]), ia.from_numpy(st["u"]), ia.from_numpy(uOld)
        u2d = ud if alias else ud.clone(memory_format=torch.preserve_format)
        assert (u2d.data_ptr() == ud.data_ptr()) == alias
        for dirO in ((3, 1, 2), (1, 2, 3)):
            stt = ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], ud, u2d, 1.0, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"],
                                    d["nhat"], uod, d["alpha"], d["drho"], st["lam_rho"], "Koren", "WH", st["uBC"], perdir, False, dirO)
            assert stt == 0
        outs.append([ia.to_numpy(x) for x in (fd, d["rhou"], d["cbar"])])
    for x, y in zip(*outs):
        assert np.array_equal(inside(x, 3), inside(y, 3))
