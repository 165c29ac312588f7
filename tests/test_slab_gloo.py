"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes run the slab decomposition with the ORACLE as the
per-rank step function and must reproduce the single-domain oracle result bit for bit on every owned cell.
This pins the overlap width W, the exchange routine and the slab geometry without a GPU."""
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


def _worker(rank, world, port, N, per_z, nsteps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import interfaceadvection.jl_b200 as ia
        from interfaceadvection.jl_b200 import configs, slab

        perdir = (1, 2, 3) if per_z else (1, 2)
        N1, N2, nz = N
        g = slab.SlabGeom(rank, world, nz, slab.W_DEFAULT, per_z)
        Ng_glob = (N1, N2, nz * world)
        Nl = (N1, N2, g.nz_local)
        lperdir = g.local_perdir(perdir)
        T = np.float64
        # global fields (every rank builds them: numpy's SIMD sin/cos is not bit-reproducible across array offsets, so the
        # local inputs are SLICED from the global arrays instead of being regenerated)
        fg0 = O.zeros(tuple(n + 2 for n in Ng_glob), T); ag = O.zeros(fg0.shape, T); ng = O.zeros(fg0.shape + (3,), T)
        sdf = configs.sdf_sphere([N1 / 2, N2 / 2, nz * world / 2], min(N1, N2) / 3.2)
        O.applyVOF(fg0, ag, ng, sdf); O.BCf(fg0, perdir)
        ug0 = np.asfortranarray(configs.tgv(Ng_glob, T, U=0.3)); O.BC(ug0, (0, 0, 0), False, perdir)
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
        for n in range(nsteps):
            dirO = tuple((1 + n + i) % 3 + 1 for i in range(1, 4))
            _oracle_step(f, u, 1e-3, lperdir, dirO)
            slab.exchange_overlap([f_t], g)
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
    N = (12, 10, 18)  # per rank: 18 owned planes (> 2W)
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

    g = SlabGeom(0, 4, 128, 8, False)
    assert (g.wlo, g.whi, g.nz_local, g.z_origin, g.lower, g.upper) == (0, 8, 136, 0, None, 1)
    g = SlabGeom(3, 4, 128, 8, False)
    assert (g.wlo, g.whi, g.nz_local, g.z_origin, g.lower, g.upper) == (8, 0, 136, 376, 2, None)
    g = SlabGeom(0, 4, 128, 8, True)
    assert (g.wlo, g.whi, g.z_origin, g.lower, g.upper) == (8, 8, -8, 3, 1)
    assert g.local_perdir((1, 2, 3)) == (1, 2)
    g = SlabGeom(0, 1, 128, 8, True)
    assert (g.wlo, g.whi, g.lower, g.upper, g.local_perdir((1, 2, 3))) == (0, 0, None, None, (1, 2, 3))
