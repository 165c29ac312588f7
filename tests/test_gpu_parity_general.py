"""Parity of the sm_100a path against the CPU oracle on the inputs REAL simulations produce (round-2 additions):

  * advectVOFρuu! with three distinct velocity arrays u¹ ≠ u² ≠ uOld (the corrector of src/flow.jl:92 has uOld = uⁿ ≠ u = uⁿ⁺½;
    the general, non-SAMEU kernel instantiations);
  * advect! / advectVOF! with u⁰ ≠ u (src/advection.jl:17, time-varying prescribed velocity);
  * a two-step MPFMomStep! in which a `project` hook changes u between predictor and corrector (src/flow.jl:75-82,95-106);
  * every limiter in Float32 as well as Float64;
  * one-step parity AT THE NAMED SIZES of BASELINE.json's configs C2 (256³, f32 + f64), C3 (512x256x256 f32), C4 (512³ f32)
    against the OpenMP build of the oracle;
  * total mass over 1000 steps (north_star) and over one full Zalesak revolution (2048 steps, config C1).

Tolerances are the north star's: max|Δf|, |Δρu| <= 1e-12 (Float64), <= 1e-5 (Float32) after one step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as O  # noqa: E402
from tests.helpers import (alloc_cmom, dirO_for, inside, make_state, oracle_cmom_call, oracle_mom_step_hook,  # noqa: E402
                           second_velocity)

TOL = {np.float32: 1e-5, np.float64: 1e-12}


@pytest.fixture(scope="module")
def ia():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    import interfaceadvection.jl_b200 as ia

    return ia


def _dev(ia, a):
    return {k: ia.from_numpy(v) for k, v in a.items()}


def _run_cuda_cmom(ia, st, f, u1, u2, uOld, rhou, dt, dirO, lam="Koren", scheme="WH"):
    a = alloc_cmom(st)
    a["rhou"][...] = rhou
    d = _dev(ia, a)
    fd, u1d, u2d, uod = ia.from_numpy(f), ia.from_numpy(u1), ia.from_numpy(u2), ia.from_numpy(uOld)
    assert len({u1d.data_ptr(), u2d.data_ptr(), uod.data_ptr()}) == 3  # three distinct device arrays: the general kernels run
    status = ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], u1d, u2d, dt, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"],
                               d["nhat"], uod, d["alpha"], d["drho"], st["lam_rho"], lam, scheme, st["uBC"], st["perdir"], False, dirO)
    return status, ia.to_numpy(fd), ia.to_numpy(d["rhou"]), ia.to_numpy(d["cbar"])


GENERAL_CASES = [
    # N, kind, perdir, uBC
    ((24, 16), "C1", (), (0, 0)),
    ((24, 16), "C3", (1, 2), (0, 0)),
    ((40, 24, 20), "C3", (), (0, 0, 0)),            # walls everywhere: boundary instantiations of every sweep
    ((36, 20, 18), "C4", (1, 2, 3), (0, 0, 0)),     # fully periodic
    ((70, 20, 12), "C2", (2,), (0, 0, 0)),          # several x tiles, mixed
    ((33, 17, 19), "C3", (1, 3), (0, 0, 0)),        # ragged extents
    ((28, 18, 16), "C2", (2,), (0.1, 0.0, 0.0)),    # inflow Dirichlet planes
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir,uBC", GENERAL_CASES)
def test_cmom_three_distinct_velocity_arrays(ia, T, N, kind, perdir, uBC):
    """advectVOFρuu! with u¹ ≠ u² ≠ uOld in VALUE (not only in address): the flux velocity is their average (advection.jl:110), the
    dilation uses both divergences (flow.jl:216) and the dilation source multiplies uOld (flow.jl:229)."""
    D = len(N)
    st = make_state(N, kind, T, perdir=perdir, uBC=uBC, scale_u=0.7)
    u1 = st["u"]
    u2 = second_velocity(st, 1, 0.6, 0.03)
    uOld = second_velocity(st, 2, 0.9, 0.02)
    assert np.abs(u1 - u2).max() > 1e-2 and np.abs(u1 - uOld).max() > 1e-3
    # ρu is what the caller built from uOld (flow.jl:91)
    a0 = alloc_cmom(st)
    O.u2rhou(a0["rhou"], uOld, st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
    rhou0 = a0["rhou"].copy(order="F")
    for dirO in ([(1, 2), (2, 1)] if D == 2 else [(3, 1, 2), (1, 2, 3), (2, 3, 1)]):
        f_o = st["f"].copy(order="F")
        so, rep, ao = oracle_cmom_call(st, f_o, u1, u2, uOld, rhou0, 1.0, dirO)
        sc, f_c, ru_c, cb_c = _run_cuda_cmom(ia, st, st["f"], u1, u2, uOld, rhou0, 1.0, dirO)
        assert sc == so, (dirO, sc, so)
        assert np.abs(f_c - f_o).max() <= TOL[T], (dirO, np.abs(f_c - f_o).max())
        scale = max(1.0, np.abs(inside(ao["rhou"], D)).max())
        err = np.abs(inside(ru_c, D) - inside(ao["rhou"], D)).max()
        assert err <= TOL[T] * scale, (dirO, err)
        assert np.array_equal(inside(cb_c, D), inside(ao["cbar"], D))
        g = f_c.copy(order="F"); O.BCf(g, st["perdir"])
        assert np.array_equal(g, f_c)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir", [((32, 32), "C1", ()), ((30, 20, 16), "C2", ()), ((24, 20, 18), "C2", (1, 2, 3)),
                                           ((70, 12, 10), "C3", (3,))])
def test_advect_with_distinct_u0_and_u(ia, T, N, kind, perdir):
    """advect!(a,c) with a time-varying prescribed velocity: u⁰ ≠ u in value (advection.jl:17-23; the flux uses their average)."""
    D = len(N)
    st = make_state(N, kind, T, perdir=perdir)
    u0 = st["u"]
    u = second_velocity(st, 3, 0.5, 0.04)
    for n in range(D):
        dirO = dirO_for(n, D)
        a = alloc_cmom(st)
        f_o = st["f"].copy(order="F")
        so, rep = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], u0, u, 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", perdir, dirO)
        d = _dev(ia, alloc_cmom(st))
        fd, u0d, ud = ia.from_numpy(st["f"]), ia.from_numpy(u0), ia.from_numpy(u)
        sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], u0d, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", perdir, dirO)
        assert sc == so
        assert np.abs(ia.to_numpy(fd) - f_o).max() <= TOL[T], dirO
        ruf_c, ruf_o = ia.to_numpy(d["rhouf"]), a["rhouf"]
        for j in range(D):
            sl = [slice(1, -1)] * D; sl[j] = slice(1, None)
            assert np.abs(ruf_c[tuple(sl) + (j,)] - ruf_o[tuple(sl) + (j,)]).max() <= TOL[T], (dirO, j)


def _hook_numpy(st):
    """A deterministic stand-in for forcing + projection: u <- BC!(0.85 u + 0.15 ρu2u(ρu,f) + smooth push)."""
    def hook(u, rhou, f, stage):
        T = st["dtype"]
        v = O.zeros(u.shape, T)
        O.rhou2u(v, rhou, f, st["lam_rho"])
        D = st["D"]
        sl = tuple([slice(1, -1)] * D)
        u[sl] = T(0.85) * u[sl] + T(0.15) * v[sl]
        O.BC(u, st["uBC"], False, st["perdir"])
    return hook


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir", [((32, 24, 20), "C3", ()), ((32, 24, 20), "C4", (1, 2)), ((24, 16), "C3", (1, 2))])
def test_two_steps_with_projection_hook(ia, T, N, kind, perdir):
    """MPFMomStep! twice with u CHANGED between predictor and corrector by a `project` hook (flow.jl:75-82,95-106 stand-in), so the
    corrector runs with uOld = uⁿ ≠ u = uⁿ⁺½ and ρu rebuilt from uⁿ -- against the oracle doing exactly the same."""
    D = len(N)
    TT = getattr(torch, np.dtype(T).name)
    st = make_state(N, kind, T, perdir=perdir, scale_u=0.8)
    hook_np = _hook_numpy(st)
    # oracle
    f_o = st["f"].copy(order="F"); u_o = st["u"].copy(order="F")
    ru_o = None
    for n in range(2):
        ru_o = oracle_mom_step_hook(st, f_o, u_o, 1.0, dirO_for(n, D), hook_np)
    # device
    flow = ia.Flow(st["N"], st["uBC"], T=TT, dt=1.0, perdir=st["perdir"])
    intf = ia.cVOF(st["N"], T=TT, lam_rho=st["lam_rho"], perdir=st["perdir"])
    flow.u.copy_(ia.from_numpy(st["u"])); intf.f.copy_(ia.from_numpy(st["f"]))

    def project(a, c, stage):
        v = torch.empty_like(a.u)
        ia.rhou2u(v, c.rhou, c.f0 if stage == "predictor" else c.f, c.lam_rho)
        sl = tuple([slice(1, -1)] * D)
        a.u[sl] = a.u[sl] * 0.85 + v[sl] * 0.15
        ia.BC(a.u, a.uBC, False, a.perdir)

    for n in range(2):
        ia.mom_advect_step(flow, intf, 1.0, project=project, check=True)
        flow.dt.append(1.0)
    tol = TOL[T] * 4  # two steps, each within the one-step tolerance, plus the hook's own round-off
    assert np.abs(ia.to_numpy(intf.f) - f_o).max() <= tol
    assert np.abs(ia.to_numpy(flow.u) - u_o).max() <= tol
    assert np.abs(inside(ia.to_numpy(intf.rhou), D) - inside(ru_o, D)).max() <= tol * max(1.0, np.abs(ru_o).max())


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("lam", ["upwind", "minmod", "Koren", "vanAlbada1", "Sweby", "superbee", "TVDcen", "TVDdown", "quick", "vanLeer", "cds"])
def test_cmom_all_limiters_both_precisions(ia, T, lam):
    st = make_state((20, 14, 12), "C2", T, perdir=(1,))
    u2 = second_velocity(st, 5, 0.8, 0.02)
    a0 = alloc_cmom(st)
    O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
    rhou0 = a0["rhou"].copy(order="F")
    uOld = st["u"].copy(order="F")
    f_o = st["f"].copy(order="F")
    so, rep, ao = oracle_cmom_call(st, f_o, st["u"], u2, uOld, rhou0, 1.0, (3, 1, 2), lam=lam)
    sc, f_c, ru_c, _ = _run_cuda_cmom(ia, st, st["f"], st["u"], u2, uOld, rhou0, 1.0, (3, 1, 2), lam=lam)
    assert np.abs(f_c - f_o).max() <= TOL[T]
    assert np.abs(inside(ru_c, 3) - inside(ao["rhou"], 3)).max() <= TOL[T]


# ---- one-step parity at the NAMED sizes ---------------------------------------------------------------------------------------------
def _named_cmom(ia, name, N, kind, perdir, T):
    from tests.helpers import oracle_mom_advect_step
    TT = getattr(torch, np.dtype(T).name)
    st = make_state(N, kind, T, perdir=perdir)
    f_o = st["f"].copy(order="F")
    ru_o = oracle_mom_advect_step(st, f_o, st["u"], 1.0, (3, 1, 2), omp=True)
    flow = ia.Flow(st["N"], st["uBC"], T=TT, dt=1.0, perdir=st["perdir"])
    intf = ia.cVOF(st["N"], T=TT, lam_rho=st["lam_rho"], perdir=st["perdir"])
    flow.u.copy_(ia.from_numpy(st["u"])); intf.f.copy_(ia.from_numpy(st["f"]))
    ia.mom_advect_step(flow, intf, 1.0)
    torch.cuda.synchronize()
    f_c = ia.to_numpy(intf.f)
    ru_c = ia.to_numpy(intf.rhou)
    del flow, intf
    torch.cuda.empty_cache()
    df = np.abs(f_c - f_o).max()
    dru = np.abs(inside(ru_c, 3) - inside(ru_o, 3)).max()
    nint = int(((f_o > 0) & (f_o < 1)).sum())
    print(f"{name}: max|Δf|={df:.3e} max|Δρu|={dru:.3e} interface cells={nint}")
    assert nint > 1000
    assert df <= TOL[T], (name, df)
    assert dru <= TOL[T] * max(1.0, float(np.abs(ru_o).max())), (name, dru)


def test_named_size_C3_dambreak_512x256x256_f32(ia):
    """BASELINE config 3 at its named size: one CMOM advection step (walls, λρ = 1e-3, Koren, WH) vs the OpenMP oracle."""
    _named_cmom(ia, "C3 512x256x256 f32", (512, 256, 256), "C3", (), np.float32)


def test_named_size_C4_bubble_512_f32(ia):
    """BASELINE config 4 at its named size (the bench workload): one CMOM advection step vs the OpenMP oracle."""
    _named_cmom(ia, "C4 512^3 f32", (512, 512, 512), "C4", (1, 2), np.float32)


def test_named_size_C4_bubble_256_f64(ia):
    """The Float64 bench line's grid (256³, periodic x/y): one CMOM advection step vs the OpenMP oracle at 1e-12."""
    _named_cmom(ia, "C4 256^3 f64", (256, 256, 256), "C4", (1, 2), np.float64)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_named_size_C2_enright_256(ia, T):
    """BASELINE config 2 at its named size: one advect! (three sweeps, discrete-curl LeVeque field) vs the OpenMP oracle."""
    N = (256, 256, 256)
    st = make_state(N, "C2", T)
    a = alloc_cmom(st)
    f_o = st["f"].copy(order="F")
    dirO = dirO_for(0, 3)
    so, rep = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], st["u"], st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", (), dirO,
                          omp=True)
    d = _dev(ia, dict(ff=a["ff"], alpha=a["alpha"], nhat=a["nhat"], cbar=a["cbar"], rhouf=a["rhouf"]))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), dirO)
    assert sc == so
    df = np.abs(ia.to_numpy(fd) - f_o).max()
    print(f"C2 256^3 {np.dtype(T).name}: max|Δf|={df:.3e}")
    assert df <= TOL[T]
    del d, fd, ud
    torch.cuda.empty_cache()


# ---- long-run mass conservation --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,bound", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_mass_1000_steps_enright_64(ia, T, bound):
    """north_star: "total mass conserved to machine precision over 1000 steps".  Time-reversed LeVeque/Enright deformation on 64³
    (discrete curl of a vector potential, exactly solenoidal), pure VOF advect!, 1000 steps: Σf drifts only by round-off and
    the 10·eps wisp snapping.  Float32 bound: 3000 sweeps x eps_32 accumulation on ≈3.7e3 cells of volume."""
    from interfaceadvection.jl_b200 import configs
    N = (64, 64, 64)
    TT = getattr(torch, np.dtype(T).name)
    dev = torch.device("cuda", 0)
    case = configs.make_case(N, dtype=np.dtype(T).name, device=dev, kind="C2")
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 64.0, T=TT, InterfaceSDF=case["sdf"], perdir=(), U=1.0, dt=1.0, device=dev)
    ubase = case["u"].clone(memory_format=torch.preserve_format)
    ia.BC(ubase, (0, 0, 0), False, ())
    V0 = ia.sum_inside(sim.intf.f)
    worst = 0.0
    nsteps = 1000
    for n in range(nsteps):
        # u⁰ = u(tⁿ), u = u(tⁿ⁺¹): both scalings of one solenoidal field, so their average is solenoidal too
        sim.flow.u0.copy_(ubase * float(np.cos(np.pi * n / nsteps)))
        sim.flow.u.copy_(ubase * float(np.cos(np.pi * (n + 1) / nsteps)))
        st = ia.advect(sim.flow, sim.intf, check=(n % 100 == 99))
        sim.flow.dt.append(1.0)
        assert st == 0
        if n % 50 == 49:
            worst = max(worst, abs(ia.sum_inside(sim.intf.f) - V0) / V0)
    V1 = ia.sum_inside(sim.intf.f)
    worst = max(worst, abs(V1 - V0) / V0)
    print(f"Enright 64^3 {np.dtype(T).name}: 1000 steps, worst relative mass drift {worst:.3e}")
    assert worst <= bound, worst
    assert float(sim.intf.f.min()) >= 0.0 and float(sim.intf.f.max()) <= 1.0


@pytest.mark.parametrize("T,bound", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_mass_1000_cmom_steps_periodic_tgv_48(ia, T, bound):
    """The same requirement on the FULL CMOM path: 1000 transport steps of MPFMomStep! (2000 advectfq! calls, 6000 sweeps) on a
    periodic 48³ Taylor-Green droplet (discretely solenoidal, test/alloctest.jl:16-21); Σf conserved to round-off, f ∈ [0,1], ρu finite."""
    from interfaceadvection.jl_b200 import configs
    N = (48, 48, 48)
    TT = getattr(torch, np.dtype(T).name)
    per = (1, 2, 3)
    sim = ia.TwoPhaseSimulation(N, (0, 0, 0), 48.0, T=TT, lam_rho=1e-3, InterfaceSDF=configs.sdf_sphere([24, 24, 24], 12.0), perdir=per,
                                U=1.0, dt=1.0)
    sim.flow.u.copy_(ia.from_numpy(configs.tgv(N, T, U=0.25))); ia.BC(sim.flow.u, (0, 0, 0), False, per)
    V0 = ia.sum_inside(sim.intf.f)
    for n in range(1000):
        ia.mom_advect_step(sim.flow, sim.intf, 1.0, check=(n % 100 == 99))
        sim.flow.dt.append(1.0)
    V1 = ia.sum_inside(sim.intf.f)
    drift = abs(V1 - V0) / V0
    print(f"TGV droplet 48^3 {np.dtype(T).name}: 1000 CMOM steps, relative mass drift {drift:.3e}")
    assert drift <= bound, drift
    assert float(sim.intf.f.min()) >= 0.0 and float(sim.intf.f.max()) <= 1.0
    assert bool(torch.isfinite(sim.intf.rhou[1:-1, 1:-1, 1:-1]).all())


def test_zalesak_full_revolution_2048_steps(ia):
    """BASELINE config 1 end to end: 128² Float64 slotted disk, Ω = 2π/2048, Δt = 1, 2048 steps = one revolution.
    Σf conserved to 1e-12·V₀; the disk returns to its place (L1 shape error of a PLIC scheme at this resolution: a few % of V₀)."""
    T = np.float64
    st = make_state((128, 128), "C1", T)
    d = _dev(ia, alloc_cmom(st))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    f_init = fd.clone(memory_format=torch.preserve_format)
    V0 = ia.sum_inside(fd)
    nsteps = 2048
    for n in range(nsteps):
        sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), dirO_for(n, 2),
                          check=(n % 256 == 255), want_rhouf=False)
        assert sc == 0
    V1 = ia.sum_inside(fd)
    l1 = float((fd - f_init)[1:-1, 1:-1].abs().sum()) / V0
    print(f"Zalesak 128^2 f64: 2048 steps, relative mass drift {abs(V1 - V0) / V0:.3e}, L1 shape error / V0 = {l1:.4f}")
    assert abs(V1 - V0) <= 1e-12 * V0
    assert l1 < 0.08
    assert float(fd.min()) >= 0.0 and float(fd.max()) <= 1.0


# ---- error behaviour (reportFillError, src/advection.jl:145-189) -----------------------------------------------------------------
def _explosive_case(T, amp):
    st = make_state((16, 16), "C1", T)
    u = np.asfortranarray(st["u"] * T(0))
    u[8, 8, 0] = T(amp); u[9, 8, 0] = T(-amp)   # a strongly converging pair of faces around cell (9,9)
    f = st["f"].copy(order="F"); f[...] = T(0.4)  # c̄ = 0: no dilation correction, the cell over-fills
    return st, f, u


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_divergence_explosion_is_fatal(ia, T):
    """|∇·u⁰|+|∇·u| > 10 at the over-filled cell: the reference throws "divergence, …, is exploding!" (advection.jl:160); a milder
    over-fill is only printed (status bits).  Both against the oracle's restatement of the same branch."""
    st, f, u = _explosive_case(T, 6.0)
    a = alloc_cmom(st)
    f_o = f.copy(order="F")
    so, rep_o = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], u, u, 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", (), (1, 2))
    assert so == -5 and rep_o.div_u0 + rep_o.div_u == pytest.approx(24.0)
    d = _dev(ia, alloc_cmom(st))
    fd, ud = ia.from_numpy(f), ia.from_numpy(u)
    with pytest.raises(ia.IfadvError, match="exploding"):
        ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2))
    # below the threshold: advisory status with the same bits, cell and divergences as the oracle
    st, f, u = _explosive_case(T, 2.0)
    a = alloc_cmom(st)
    f_o = f.copy(order="F")
    so, rep_o = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], u, u, 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", (), (1, 2))
    d = _dev(ia, alloc_cmom(st))
    fd, ud = ia.from_numpy(f), ia.from_numpy(u)
    rep = ia.Report()
    sc = ia.context_for(fd).advect_vof(torch.cuda.current_stream().cuda_stream, fd.data_ptr(), d["ff"].data_ptr(), d["alpha"].data_ptr(),
                                       d["nhat"].data_ptr(), ud.data_ptr(), ud.data_ptr(), 1.0, d["cbar"].data_ptr(), d["rhouf"].data_ptr(),
                                       st["lam_rho"], 0, (), (1, 2), 0, rep)
    assert sc == so and sc > 0
    assert rep.maxf == pytest.approx(rep_o.maxf, rel=1e-6) and tuple(rep.argmax)[:2] == tuple(rep_o.argmax)[:2]
    assert np.abs(ia.to_numpy(fd) - f_o).max() <= TOL[T]


def test_sticky_nan_without_report(ia):
    """Calls made without a report never synchronise; a NaN they produce is remembered on the device and surfaces at the next
    ifadv_check_nan / reporting call (error("NaN!"), advection.jl:148)."""
    st = make_state((12, 12), "C1", np.float64)
    bad = st["f"].copy(order="F"); bad[5, 5] = np.nan
    d = _dev(ia, alloc_cmom(st))
    fd, ud = ia.from_numpy(bad), ia.from_numpy(st["u"])
    ctx = ia.context_for(fd)
    s = torch.cuda.current_stream().cuda_stream
    try:
        ctx.check_nan(s)  # clean slate: the context is shared (cached by shape) with other tests that leave a NaN behind
    except ia.IfadvError:
        pass
    ctx.check_nan(s)      # reading the flags consumed them
    assert ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2), check=False) == 0
    # a later, healthy call WITH a report still reports the earlier NaN
    gd = ia.from_numpy(st["f"])
    with pytest.raises(ia.IfadvError, match="NaN"):
        ia.advectVOF(gd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2), check=True)
    # the flag is one-shot
    assert ia.advectVOF(gd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2), check=True) == 0
    # and ifadv_check_nan sees a NaN of the most recent unreported call directly
    fd2 = ia.from_numpy(bad)
    ia.advectVOF(fd2, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2), check=False)
    with pytest.raises(ia.IfadvError, match="NaN"):
        ctx.check_nan(s)
    ia.advectVOF(gd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 2), check=False)
    ctx.check_nan(s)


def test_rejected_calls_leave_the_context_clean(ia):
    """dirO must be a permutation of 1..D (a repeated direction used to run silently); a rejected CMOM call consumes the one-shot
    event of ifadv_defer_f_writes_until, so the next call on the shared context does not wait on a stale handle."""
    st = make_state((12, 10, 8), "C3", np.float64)
    d = _dev(ia, alloc_cmom(st))
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    with pytest.raises(ia.IfadvError, match="permutation"):
        ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", (), (1, 1, 2))
    ctx = ia.context_for(fd)
    ev = torch.cuda.Event()
    ev.record()
    ctx.defer_f_writes_until(ev.cuda_event)
    with pytest.raises(ia.IfadvError):  # rejected: invalid limiter
        ctx.advect_vof_rhouu(0, fd.data_ptr(), d["ff"].data_ptr(), d["alpha"].data_ptr(), d["nhat"].data_ptr(), ud.data_ptr(), ud.data_ptr(), 1.0,
                             d["cbar"].data_ptr(), d["rhou"].data_ptr(), d["r"].data_ptr(), d["Phi"].data_ptr(), d["rhouf"].data_ptr(),
                             d["nhat"].data_ptr(), ud.data_ptr(), d["alpha"].data_ptr(), d["drho"].data_ptr(), 1e-3, 99, 0, (0, 0, 0), (), False,
                             (3, 1, 2))
    del ev  # the event is gone; a stale handle in the context would now fail with -3
    torch.cuda.synchronize()
    assert ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"], d["nhat"], ud,
                             d["alpha"], d["drho"], 1e-3, "Koren", "WH", (0, 0, 0), (), False, (3, 1, 2)) == 0


# ---- exitBC = true (BC!'s saveexit, flow.jl:197,207) -----------------------------------------------------------------------------
def _exit_case(N, kind, T, perdir):
    """Inflow uBC[1] through x = 1, 2 and a free exit plane x = N: u[N,·,1] differs from uBC, the caller's r array holds an arbitrary
    'saved' exit value of u★ on plane N of component 1 and garbage everywhere else (only that plane may be read)."""
    D = len(N)
    uBC = (0.2,) + (0.0,) * (D - 1)
    st = make_state(N, kind, T, perdir=perdir, uBC=uBC, scale_u=0.5)
    rng = np.random.default_rng(20261018)
    u = st["u"]
    u[-1, ..., 0] = (0.2 + 0.05 * rng.standard_normal(u[-1, ..., 0].shape)).astype(T)  # exit plane: outflow, not the Dirichlet value
    O.BC(u, uBC, True, perdir)
    a = alloc_cmom(st)
    a["r"][...] = rng.standard_normal(a["r"].shape).astype(T)
    a["r"][-1, ..., 0] = (0.2 + 0.05 * rng.standard_normal(a["r"][-1, ..., 0].shape)).astype(T)
    O.u2rhou(a["rhou"], u, st["f"], st["lam_rho"]); O.BC(a["rhou"], uBC, True, perdir)
    return st, a


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir", [((24, 16), "C1", (2,)), ((20, 12), "C3", ()), ((24, 14, 10), "C2", (2, 3)), ((70, 12, 9), "C3", ())])
def test_cmom_exit_bc_matches_oracle(ia, T, N, kind, perdir):
    """advectVOFρuu! with exitBC=true.  The reference never defines plane N of component 1 of u★: BC! with saveexit keeps whatever
    r[N,·,1] holds (flow.jl:197) -- the caller's value in the first sweep, flux residue of the previous advectρuu1D! afterwards.  The B200
    path reads the caller's value in every sweep, so it equals the oracle bit for bit (f64) when x is swept first, and
    everywhere except the last interior x-plane of ρu[:,1] for the other orders.  The exit face keeps its computed mass flux."""
    D = len(N)
    orders = [(1, 2), (2, 1)] if D == 2 else [(1, 2, 3), (3, 1, 2), (2, 3, 1)]
    for dirO in orders:
        st, a = _exit_case(N, kind, T, perdir)
        ao = {k: v.copy(order="F") for k, v in a.items()}
        f_o = st["f"].copy(order="F")
        so, rep = O.advectVOFrhouu(f_o, ao["ff"], ao["alpha"], ao["nhat"], st["u"], st["u"], 1.0, ao["cbar"], ao["rhou"], ao["r"], ao["Phi"],
                                   ao["rhouf"], ao["nhat"], st["u"], ao["alpha"], ao["drho"], st["lam_rho"], "Koren", "WH", st["uBC"], perdir,
                                   True, dirO)
        d = _dev(ia, a)
        fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
        sc = ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"], d["nhat"],
                               ud, d["alpha"], d["drho"], st["lam_rho"], "Koren", "WH", st["uBC"], perdir, True, dirO)
        assert sc == so
        assert np.abs(ia.to_numpy(fd) - f_o).max() <= TOL[T], dirO
        ru_c, ru_o = inside(ia.to_numpy(d["rhou"]), D), inside(ao["rhou"], D)
        err = np.abs(ru_c - ru_o)
        if dirO[0] != 1:
            err[-1, ..., 0] = 0  # the reference's u★[N,·,1] is unspecified scratch there (see the docstring)
        scale = max(1.0, np.abs(ru_o).max())
        assert err.max() <= TOL[T] * scale, (dirO, err.max(), np.unravel_index(err.argmax(), err.shape))
        # the exit really is free: the result differs from the Dirichlet treatment of plane N
        if dirO[0] == 1:
            st2, a2 = _exit_case(N, kind, T, perdir)
            d2 = _dev(ia, a2)
            fd2 = ia.from_numpy(st2["f"])
            ia.advectVOFrhouu(fd2, d2["ff"], d2["alpha"], d2["nhat"], ud, ud, 1.0, d2["cbar"], d2["rhou"], d2["r"], d2["Phi"], d2["rhouf"],
                              d2["nhat"], ud, d2["alpha"], d2["drho"], st["lam_rho"], "Koren", "WH", st["uBC"], perdir, False, dirO)
            assert np.abs(inside(ia.to_numpy(d2["rhou"]), D)[-1, ..., 0] - ru_c[-1, ..., 0]).max() > 1e-4


def test_exit_bc_through_the_fused_entry_and_the_mirror(ia):
    """The fused entry (u2ρu! + BC!(…, exitBC) + advectVOFρuu!) equals the three separate calls with exitBC=true (Float64: bit for bit)."""
    T, N, perdir = np.float64, (70, 12, 9), ()
    st, a = _exit_case(N, "C3", T, perdir)
    dirO = (1, 2, 3)
    d = _dev(ia, a)
    fd, ud = ia.from_numpy(st["f"]), ia.from_numpy(st["u"])
    ia.advectVOFrhouu(fd, d["ff"], d["alpha"], d["nhat"], ud, ud, 1.0, d["cbar"], d["rhou"], d["r"], d["Phi"], d["rhouf"], d["nhat"], ud,
                      d["alpha"], d["drho"], st["lam_rho"], "Koren", "WH", st["uBC"], perdir, True, dirO)
    st2, a2 = _exit_case(N, "C3", T, perdir)
    d2 = _dev(ia, a2)
    f2, u2 = ia.from_numpy(st2["f"]), ia.from_numpy(st2["u"])
    ctx = ia.context_for(f2)
    p = lambda t: t.data_ptr()
    ctx.u2rhou_advect_vof_rhouu(0, p(f2), p(f2), p(d2["ff"]), p(d2["Phi"]), p(u2), p(u2), 1.0, p(d2["cbar"]), p(d2["rhou"]), p(d2["r"]),
                                p(d2["rhouf"]), p(u2), p(d2["drho"]), st["lam_rho"], 2, 0, st["uBC"], perdir, True, dirO)
    torch.cuda.synchronize()
    assert np.array_equal(ia.to_numpy(f2), ia.to_numpy(fd))
    assert np.array_equal(inside(ia.to_numpy(d2["rhou"]), 3), inside(ia.to_numpy(d["rhou"]), 3))


# ---- pure VOF, 3-D: the cell-parallel kernel (ifadv_vofcell.cuh, default) against the marching generation and the oracle ---------
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir,dirO", [((70, 40, 36), "C2", (), (1, 2, 3)), ((33, 18, 21), "C2", (1, 2, 3), (3, 1, 2)),
                                                 ((40, 24, 20), "C3", (2,), (2, 3, 1)), ((36, 20, 12), "C4", (1, 3), (1, 2, 3))])
def test_pure_vof_cell_kernel_matches_oracle_and_marching_kernels(ia, T, N, kind, perdir, dirO):
    """advectVOF! with u⁰ ≠ u through both 3-D pure-VOF kernel generations: each within the north star's tolerance of the oracle
    (f, ρuf on inside_uWB faces, c̄, status) and bit-identical to each other in Float64."""
    import os
    from interfaceadvection.jl_b200 import api
    st = make_state(N, kind, T, perdir=perdir)
    u0 = second_velocity(st, 3, 0.8, 0.05)
    a = alloc_cmom(st)
    f_o = st["f"].copy(order="F")
    so, rep = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], u0, st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", perdir, dirO)
    res = {}
    try:
        for gen in (None, "lean"):
            api._contexts.clear()
            if gen is None:
                os.environ.pop("IFADV_VOF_KERNEL", None)
            else:
                os.environ["IFADV_VOF_KERNEL"] = gen
            d = _dev(ia, alloc_cmom(st))
            fd, ud, u0d = ia.from_numpy(st["f"]), ia.from_numpy(st["u"]), ia.from_numpy(u0)
            sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], u0d, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", perdir, dirO)
            res[gen] = (sc, ia.to_numpy(fd), ia.to_numpy(d["rhouf"]), ia.to_numpy(d["cbar"]))
    finally:
        os.environ.pop("IFADV_VOF_KERNEL", None)
        api._contexts.clear()
    for gen, (sc, f_c, ruf_c, cb_c) in res.items():
        assert sc == so, gen
        assert np.abs(f_c - f_o).max() <= TOL[T], gen
        for j in range(3):
            sl = [slice(1, -1)] * 3; sl[j] = slice(1, None)
            assert np.abs(ruf_c[tuple(sl) + (j,)] - a["rhouf"][tuple(sl) + (j,)]).max() <= TOL[T], (gen, j)
        assert np.array_equal(inside(cb_c, 3), inside(a["cbar"], 3)), gen
    if T == np.float64:
        assert np.array_equal(res[None][1], res["lean"][1])
        for j in range(3):
            sl = [slice(1, -1)] * 3; sl[j] = slice(1, None)
            assert np.array_equal(res[None][2][tuple(sl) + (j,)], res["lean"][2][tuple(sl) + (j,)])


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("N,kind,perdir", [((48, 40), "C1", ()), ((37, 29), "C3", (1,)), ((64, 24), "C1", (1, 2))])
def test_pure_vof_2d_both_fixup_variants_match_oracle(ia, T, N, kind, perdir, monkeypatch):
    """2-D advectVOF! is one cooperative launch of the cell-parallel sweeps (ifadv_vofcell.cuh): small grids reconstruct a flagged cell in
    line, large ones go through the deferred list (forced here with IFADV_VOF2D_LIST).  Both against the oracle, and bit-identical to
    each other and to the v1 tile kernel in Float64."""
    from interfaceadvection.jl_b200 import api
    st = make_state(N, kind, T, perdir=perdir)
    u0 = second_velocity(st, 5, 0.9, 0.04)
    a = alloc_cmom(st)
    f_o = st["f"].copy(order="F")
    res = {}
    for dirO in [(1, 2), (2, 1)]:
        f_o = st["f"].copy(order="F"); a = alloc_cmom(st)
        so, rep = O.advectVOF(f_o, a["ff"], a["alpha"], a["nhat"], u0, st["u"], 1.0, a["cbar"], a["rhouf"], st["lam_rho"], "WH", perdir, dirO)
        for var in ("local", "list", "tile"):
            monkeypatch.delenv("IFADV_VOF2D_LIST", raising=False); monkeypatch.delenv("IFADV_KERNEL", raising=False)
            if var == "list":
                monkeypatch.setenv("IFADV_VOF2D_LIST", "1")
            if var == "tile":
                monkeypatch.setenv("IFADV_KERNEL", "tile")
            api._contexts.clear()
            d = _dev(ia, alloc_cmom(st))
            fd, ud, u0d = ia.from_numpy(st["f"]), ia.from_numpy(st["u"]), ia.from_numpy(u0)
            sc = ia.advectVOF(fd, d["ff"], d["alpha"], d["nhat"], u0d, ud, 1.0, d["cbar"], d["rhouf"], st["lam_rho"], "WH", perdir, dirO)
            res[var] = (ia.to_numpy(fd), ia.to_numpy(d["rhouf"]))
            assert sc == so, (var, dirO)
            assert np.abs(res[var][0] - f_o).max() <= TOL[T], (var, dirO)
            for j in range(2):
                sl = [slice(1, -1)] * 2; sl[j] = slice(1, None)
                assert np.abs(res[var][1][tuple(sl) + (j,)] - a["rhouf"][tuple(sl) + (j,)]).max() <= TOL[T], (var, dirO, j)
            assert np.array_equal(inside(ia.to_numpy(d["cbar"]), 2), inside(a["cbar"], 2))
        monkeypatch.delenv("IFADV_VOF2D_LIST", raising=False); monkeypatch.delenv("IFADV_KERNEL", raising=False)
        api._contexts.clear()
        assert np.array_equal(res["local"][0], res["list"][0])
        if T == np.float64:
            assert np.array_equal(res["local"][0], res["tile"][0])
