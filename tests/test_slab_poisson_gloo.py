"""N>1 host logic of the pressure solver on CPU: world_size-2 (and 3) gloo processes run the z-slab form of psolver!
(src/flow.jl:300-326) -- the scheme of ifadv_psolver on a slab context (csrc/ifadv_poisson.cu): kernels over the owned planes, the
ranks' partial sums all-reduced before every scalar step, the ghost planes of ϵ (every iteration) and x (both ends) from the
z-neighbours -- with the ORACLE's operator as the per-rank kernel, and must reproduce the single-domain oracle solve: same iteration
count, solution to round-off of the dot products.  Pins where the exchanges and all-reduces sit, the cell count of the residual mean,
and the periodic wrap through the exchange, without a GPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as O
from tests.test_slab_gloo import _free_port, _np_view


def _slab_psolver(slab, g, p, x_t, eps_t, lperdir, tol, itmx):
    """psolver! on one slab.  p: local oracle Poisson whose x / eps are numpy views of the column-major tensors x_t / eps_t."""
    T = p.x.dtype.type
    ins = (slice(1, -1), slice(1, -1), g.owned)

    def allsum(*vals):
        s = torch.tensor([float(v) for v in vals], dtype=torch.float64)
        dist.all_reduce(s)
        return [float(v) for v in s]

    def dot(a, b):
        return T(allsum((a[ins].astype(np.float64) * b[ins].astype(np.float64)).sum())[0])

    def mult(t, a):  # perBC!(a) in x, y; the z-neighbours' plane; A·a on the owned planes (into p.z)
        O.pois_mult(p, a)
        slab.exchange_overlap([t], g)
        return O.pois_mult(p, a)

    src = p.z.copy(order="F")
    Ax = mult(x_t, p.x).copy(order="F")                                    # residual!: r = z - A x, mean removed
    p.r[ins] = np.where(p.iD[ins] == 0, T(0), src[ins] - Ax[ins])
    tot, cnt = allsum(p.r[ins].astype(np.float64).sum(), p.r[ins].size)
    s = T(tot) / T(cnt)
    if abs(s) > 2 * np.finfo(T).eps:
        p.r[ins] -= s
    r2 = dot(p.r, p.r)
    p.z[ins] = p.r[ins] * p.iD[ins]; p.eps[ins] = p.z[ins]
    rho = dot(p.r, p.z)
    n = 0
    while (r2 > tol or (r2 > tol / 4 and n == 0)) and n < itmx:
        q = mult(eps_t, p.eps)                                             # q aliases p.z
        alpha = T(rho / dot(q, p.eps))
        p.x[ins] += alpha * p.eps[ins]
        p.r[ins] -= alpha * q[ins]
        p.z[ins] = p.r[ins] * p.iD[ins]
        rho2 = dot(p.r, p.z)
        beta = T(rho2 / rho)
        p.eps[ins] = beta * p.eps[ins] + p.z[ins]
        rho = rho2
        r2 = dot(p.r, p.r)
        n += 1
    O.pois_mult(p, p.x); slab.exchange_overlap([x_t], g)                    # perBC!(x) + ghost planes
    return n, float(r2)


def _worker(rank, world, port, N, per_z, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import interfaceadvection.jl_b200 as ia
        from interfaceadvection.jl_b200 import slab
        from tests.test_oracle_poisson import make_L

        T = np.float64
        perdir = (1, 3) if per_z else (2,)
        N1, N2, nz = N
        g = slab.SlabGeom(rank, world, nz, slab.G_DEFAULT, per_z)
        nzg = nz * world
        Ngg = (N1 + 2, N2 + 2, nzg + 2)
        lperdir = g.local_perdir(perdir)
        Lg = make_L(Ngg, perdir, T, seed=5)                                # global problem, identical on every rank
        rng = np.random.default_rng(6)
        b = rng.standard_normal((N1, N2, nzg)); b -= b.mean()
        zg = O.zeros(Ngg, T); zg[1:-1, 1:-1, 1:-1] = b
        zidx = [((g.z_origin + l - 1) % nzg) + 1 if per_z else min(max(g.z_origin + l, 0), nzg + 1) for l in range(g.nz_local + 2)]
        Ngl = (N1 + 2, N2 + 2, g.nz_local + 2)
        x_t = ia.jl_zeros(Ngl, torch.float64, "cpu"); eps_t = ia.jl_zeros(Ngl, torch.float64, "cpu")
        p = O.Poisson(_np_view(x_t), np.asfortranarray(Lg[:, :, zidx, :]), np.asfortranarray(zg[:, :, zidx]), lperdir)
        p.eps = _np_view(eps_t)
        tol = 50 * np.finfo(T).eps
        n, r2 = _slab_psolver(slab, g, p, x_t, eps_t, lperdir, tol, 2000)
        owned = p.x[1:-1, 1:-1, g.owned].copy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (owned, n))
        if rank == 0:
            pg = O.Poisson(O.zeros(Ngg, T), Lg, zg, perdir)
            n_ref, _ = O.psolver(pg)
            got = np.concatenate([a for a, _ in gathered], axis=2)
            ref = pg.x[1:-1, 1:-1, 1:-1]
            out.put((n_ref, [k for _, k in gathered], float(np.abs(ref - got).max()), float(np.abs(ref).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,per_z", [(2, False), (2, True), (3, False)])
def test_slab_psolver_matches_single_domain(world, per_z):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, (10, 8, 6), per_z, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    n_ref, ns, err, scale = out.get(timeout=5)
    assert all(abs(k - n_ref) <= 1 for k in ns) and len(set(ns)) == 1, (n_ref, ns)  # every rank takes the same decisions
    assert err <= 1e-9 * max(1.0, scale), err
