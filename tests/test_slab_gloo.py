"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes run the slab decomposition with the ORACLE as the
per-rank sweep function and must reproduce the single-domain oracle result bit for bit on every owned cell.
The scheme is the multi-GPU path's (include/ifadv.h): G = 3 ghost planes per neighbour side, exchange of f, ρu and c̄ after every
directional sweep.  This pins the ghost width (reach of one sweep), the plane ranges and posting order of the exchange and the
slab geometry without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_view(t):
    """Fortran-ordered numpy view sharing memory with a column-major CPU tensor."""
    return t.permute(*reversed(range(t.dim()))).numpy().T


def _oracle_step(f, u, lam_rho, perdir, dirO):
    from tests.helpers import oracle_mom_advect_step

    st = dict(D=3, Ng=f.shape, dtype=f.dtype.type, perdir=perdir, uBC=(0.0, 0.0, 0.0), lam_rho=lam_rho)
    oracle_mom_advect_step(st, f, u, 1.0, dirO)


def _slab_step(ia, slab, g, f_t, u, lam_rho, lperdir, dirO):
    """Transport half of MPFMomStep! on one slab, sweep by sweep, with the ghost-plane exchanges of the multi-GPU path."""
    from tests.helpers import alloc_cmom

    f = _np_view(f_t)
    T = f.dtype.type
    st = dict(D=3, Ng=f.shape, dtype=T, perdir=lperdir, uBC=(0.0, 0.0, 0.0), lam_rho=lam_rho)
    a = alloc_cmom(st)
    # exchanged arrays live in column-major torch tensors; numpy views share their memory
    ru_t = ia.jl_zeros(f.shape + (3,), f_t.dtype, "cpu"); ru = _np_view(ru_t)
    cb_t = ia.jl_zeros(f.shape, torch.int8, "cpu"); cb = _np_view(cb_t)
    f0_t = ia.jl_zeros(f.shape, f_t.dtype, "cpu"); f0 = _np_view(f0_t)
    u0 = u.copy(order="F")

    def group(ft, fa, u1, u2, uOld):
        O.u2rhou(ru, u0, fa, lam_rho); O.BC(ru, (0, 0, 0), False, lperdir)
        for iop in range(3):
            O.advectVOFrhouu_sweep(iop, fa, a["ff"], a["alpha"], a["nhat"], u1, u2, 1.0, cb, ru, a["r"], a["Phi"], a["rhouf"], a["nhat"],
                                   uOld, a["alpha"], a["drho"], lam_rho, "Koren", "WH", (0, 0, 0), lperdir, False, dirO)
            if iop < 2:
                slab.exchange_overlap([ft, ru_t] + ([cb_t] if iop == 0 else []), g)
            else:
                slab.exchange_overlap([ft], g)

    f0[...] = f
    group(f0_t, f0, u0, u, u)          # flow.jl:69-70
    f0[...] = (f0 + f) * T(0.5)        # :74
    f0[...] = f                        # :89
    group(f_t, f, u, u, u0)            # :91-92
def _worker(rank, world, port, N, per_z, nsteps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import interfaceadvection.jl_b200 as ia
        from interfaceadvection.jl_b200 import configs, slab

        perdir = (1, 2, 3) if per_z else (1, 2)
        N1, N2, nz = N
        g = slab.SlabGeom(rank, world, nz, slab.G_DEFAULT, per_z)
        Ng_glob = (N1, N2, nz * world)
        Nl = (N1, N2, g.nz_local)
        lperdir = g.local_perdir(perdir)
        T = np.float64
        # global fields (every rank builds them: numpy's SIMD sin/cos is not bit-reproducible across array offsets, so the
        # local inputs are SLICED from the global arrays instead of being regenerated)
        fg0 = O.zeros(tuple(n + 2 for n in Ng_glob), T); ag = O.zeros(fg0.shape, T); ng = O.zeros(fg0.shape + (3,), T)
        sdf = configs.sdf_sphere([N1 / 2, N2 / 2, nz * world / 2], min(N1, N2) / 3.2)
        O.applyVOF(fg0, ag, ng, sdf); O.BCf(fg0, perdir)
        ug0 = np.asfortranarray(configs.enright(Ng_glob, T, amp=0.3)); O.BC(ug0, (0, 0, 0), False, perdir)  # w != 0: z fluxes cross the slab ends
        # local array plane l (0-based, ghost at 0) <-> global array plane z_origin + l, wrapped on a periodic box
        nzg = nz * world
        zidx = [((g.z_origin + l - 1) % nzg) + 1 if per_z else min(max(g.z_origin + l, 0), nzg + 1) for l in range(g.nz_local + 2)]
        u = np.asfortranarray(ug0[:, :, zidx, :])
        O.BC(u, (0, 0, 0), False, lperdir)
        f_t = ia.jl_zeros(tuple(n + 2 for n in Nl), torch.float64, "cpu")
        f = _np_view(f_t)
        f[...] = fg0[:, :, zidx]
        O.BCf(f, lperdir)
        slab.exchange_overlap([f_t], g)
        u_t = ia.jl_zeros(u.shape, torch.float64, "cpu")
        _np_view(u_t)[...] = u
        slab.exchange_overlap([u_t], g)  # BC! treated the slab ends as walls: the neighbours' planes replace that
        u = _np_view(u_t)
        for n in range(nsteps):
            dirO = tuple((1 + n + i) % 3 + 1 for i in range(1, 4))
            _slab_step(ia, slab, g, f_t, u, 1e-3, lperdir, dirO)
        owned = f[1:-1, 1:-1, g.owned].copy()
        # rank 0 gathers and compares with the single-domain run
        gathered = [None] * world
        dist.all_gather_object(gathered, owned)
        if rank == 0:
            fg, ug = fg0, ug0
            for n in range(nsteps):
                dirO = tuple((1 + n + i) % 3 + 1 for i in range(1, 4))
                _oracle_step(fg, ug, 1e-3, perdir, dirO)
            ref = fg[1:-1, 1:-1, 1:-1]
            got = np.concatenate(gathered, axis=2)
            out.put((bool(np.array_equal(ref, got)), float(np.abs(ref - got).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,per_z", [(2, False), (2, True), (3, False)])
def test_slab_oracle_bitwise(world, per_z):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    N = (12, 10, 9)  # per rank: 9 owned planes (>= G)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, per_z, 2, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, err = out.get(timeout=5)
    assert ok, f"slab result differs from the single-domain run: max|Δf| = {err}"


def test_slab_geometry():
    from interfaceadvection.jl_b200.slab import SlabGeom

    g = SlabGeom(0, 4, 128, 8, False)  # any ghost width
    assert (g.wlo, g.whi, g.nz_local, g.z_origin, g.lower, g.upper) == (0, 8, 136, 0, None, 1)
    g = SlabGeom(3, 4, 128, 8, False)
    assert (g.wlo, g.whi, g.nz_local, g.z_origin, g.lower, g.upper) == (8, 0, 136, 376, 2, None)
    g = SlabGeom(0, 4, 128, 8, True)
    assert (g.wlo, g.whi, g.z_origin, g.lower, g.upper) == (8, 8, -8, 3, 1)
    assert g.local_perdir((1, 2, 3)) == (1, 2)
    g = SlabGeom(0, 1, 128, 8, True)
    assert (g.wlo, g.whi, g.lower, g.upper, g.local_perdir((1, 2, 3))) == (0, 0, None, None, (1, 2, 3))
