"""z-slab decomposition of the VOF + CMOM advection path across the GPUs of one box (one process per GPU).

The reference has no multi-device code; this is new functionality whose oracle is the single-GPU run on the same
global grid: owned cells must come out BIT-IDENTICAL.

Scheme (include/ifadv.h, "z-slab decomposition"): per-sweep ghost-plane exchange.  A rank's arrays hold its nz owned planes plus
G = 3 ghost planes per neighbour side (the reach of ONE directional sweep: 3 planes below / 2 above).  The library
(`ifadv_create_slab`) updates the owned planes only and, after every directional sweep, swaps the G boundary planes of what the
sweep produced (f, ρu x3, once c̄) with both neighbours -- `ncclSend/ncclRecv` in one group over NVLink -- so the next sweep reads
the neighbours' planes like any interior plane.  No redundant planes are computed (round 1 recomputed 16 overlap planes per
slab: 3 % at 512 planes per GPU, 12.5 % at the 128 planes per GPU of BASELINE config 5).  This module is only the CALLER of that
ABI: geometry, NCCL bootstrap through torch.distributed, and the initial fill of a slab.

`exchange_overlap` is the same exchange written with torch.distributed point-to-point ops; the CPU (gloo) tests use it to pin the
plane ranges and the posting order of the C implementation without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

G_DEFAULT = 3  # ghost planes per neighbour side = reach of one directional sweep
W_DEFAULT = G_DEFAULT


@dataclass
class SlabGeom:
    """Geometry of rank `rank`'s slab: the global grid is N[0] x N[1] x (nz_owned * world)."""
    rank: int
    world: int
    nz_owned: int
    W: int  # ghost planes per neighbour side
    per_z: bool

    @property
    def wlo(self) -> int:
        return self.W if (self.world > 1 and (self.rank > 0 or self.per_z)) else 0

    @property
    def whi(self) -> int:
        return self.W if (self.world > 1 and (self.rank < self.world - 1 or self.per_z)) else 0

    @property
    def nz_local(self) -> int:  # interior planes of the local array (owned + ghost planes)
        return self.nz_owned + self.wlo + self.whi

    @property
    def z_origin(self) -> int:  # global 0-based index of the first local interior plane
        return self.rank * self.nz_owned - self.wlo

    @property
    def owned(self) -> slice:  # owned planes as a slice of the local array's z axis (array ghost plane at 0)
        return slice(1 + self.wlo, 1 + self.wlo + self.nz_owned)

    @property
    def lower(self) -> Optional[int]:
        if self.world == 1:
            return None
        if self.rank > 0:
            return self.rank - 1
        return self.world - 1 if self.per_z else None

    @property
    def upper(self) -> Optional[int]:
        if self.world == 1:
            return None
        if self.rank < self.world - 1:
            return self.rank + 1
        return 0 if self.per_z else None

    def local_perdir(self, perdir: Sequence[int]) -> Tuple[int, ...]:
        """z is never periodic locally when the slab has neighbours: the wrap goes through the exchange."""
        return tuple(p for p in perdir if not (p == 3 and self.world > 1))


def zplanes(t: torch.Tensor, z0: int, z1: int) -> torch.Tensor:
    """Contiguous view of planes z0:z1 of a column-major (x,y,z) scalar field (storage order is (z,y,x))."""
    assert t.dim() == 3
    return t.permute(2, 1, 0)[z0:z1]


def exchange_overlap(fields: List[torch.Tensor], g: SlabGeom, dist=None) -> int:
    """Fill the ghost planes of every field (scalar (x,y,z) or vector (x,y,z,c), column-major) from the neighbours' owned planes.
    Returns bytes sent.  Same plane ranges and posting order as slab_exchange in csrc/ifadv_b200.cu: per component, sends go
    [up, down] and receives [from below, from above], so the pairs match when both neighbours are one peer (2 ranks, periodic z)."""
    if g.world == 1:
        return 0
    import torch.distributed as tdist

    dist = dist or tdist
    ops, sent = [], 0
    o = g.owned
    for fld in fields:
        comps = [fld] if fld.dim() == 3 else [fld[..., i] for i in range(fld.shape[-1])]
        for f in comps:
            if g.upper is not None:
                top = zplanes(f, o.stop - g.W, o.stop)
                ops.append(dist.P2POp(dist.isend, top, g.upper))
                sent += top.numel() * top.element_size()
            if g.lower is not None:
                bot = zplanes(f, o.start, o.start + g.W)
                ops.append(dist.P2POp(dist.isend, bot, g.lower))
                sent += bot.numel() * bot.element_size()
            if g.lower is not None and g.wlo:
                ops.append(dist.P2POp(dist.irecv, zplanes(f, o.start - g.wlo, o.start), g.lower))
            if g.upper is not None and g.whi:
                ops.append(dist.P2POp(dist.irecv, zplanes(f, o.stop, o.stop + g.whi), g.upper))
    if not ops:
        return 0
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return sent


_nccl_comms = {}


def nccl_comm(rank: int, world: int, device_index: int) -> int:
    """One NCCL communicator per process for the library's exchanges, bootstrapped through torch.distributed (rank 0 creates the
    unique id, everybody receives it).  Returns the ncclComm_t as an integer handle."""
    key = (rank, world, device_index)
    if key not in _nccl_comms:
        import torch.distributed as dist

        from . import _lib

        box = [_lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        _nccl_comms[key] = _lib.nccl_comm_init(world, box[0], rank, device_index)
    return _nccl_comms[key]


class SlabRunner:
    """One rank of the slab-decomposed CMOM advection benchmark / simulation (CUDA): builds the local arrays, binds them to an
    `ifadv_create_slab` context and steps them with the unchanged host-side mirror (mom_advect_step)."""

    def __init__(self, N_per_gpu, dtype: str, perdir, kind: str, rank: int, world: int, device, G: int = G_DEFAULT, lam_rho=None,
                 fields=None, vel=None):
        """fields: optional (f_local, u_local) column-major device tensors (e.g. slices of a global state); by default the
        slab samples the analytic generators of `kind` at its global offset."""
        import interfaceadvection.jl_b200 as ia
        from . import _lib, api, configs

        self.ia = ia
        N1, N2, nz = N_per_gpu
        self.geom = g = SlabGeom(rank, world, nz, G, 3 in tuple(perdir))
        self.Nglobal = (N1, N2, nz * world)
        self.Nlocal = (N1, N2, g.nz_local)
        self.perdir = g.local_perdir(perdir)
        T = getattr(torch, dtype)
        dev = torch.device(device)
        origin = (0, 0, g.z_origin)
        Ngl = tuple(n + 2 for n in self.Nlocal)
        if world > 1:
            self.ctx = _lib.Context(Ngl, dtype, dev.index or 0, slab=dict(comm=nccl_comm(rank, world, dev.index or 0), rank=rank,
                                                                          nranks=world, G=G, per_z=g.per_z))
            api.register_context(Ngl, T, dev.index, self.ctx)
        else:
            self.ctx = None
        self.flow = ia.Flow(self.Nlocal, (0, 0, 0), T=T, dt=1.0, perdir=self.perdir, device=device)
        if fields is None:
            case = configs.make_case(self.Nglobal, dtype=dtype, device=device, kind=kind, Nl=self.Nlocal, origin=origin, vel=vel)
            self.lam_rho = case["lam_rho"] if lam_rho is None else lam_rho
            self.intf = ia.cVOF(self.Nlocal, T=T, lam_rho=self.lam_rho, InterfaceSDF=case["sdf"], perdir=self.perdir, device=device,
                                origin=origin)
            self.flow.u.copy_(case["u"])
        else:
            self.lam_rho = 1e-3 if lam_rho is None else lam_rho
            self.intf = ia.cVOF(self.Nlocal, T=T, lam_rho=self.lam_rho, perdir=self.perdir, device=device)
            self.intf.f.copy_(fields[0])
            ia.BCf(self.intf.f, self.perdir)
            self.flow.u.copy_(fields[1])
        if self.ctx is None:
            self.ctx = ia.context_for(self.intf.f)
        ia.BC(self.flow.u, (0, 0, 0), False, self.perdir)
        self.exchange(self.flow.u)      # BC! treated the slab ends as walls: the neighbours' planes replace that
        self.exchange(self.intf.f)      # the analytic SDF filled the ghost planes already; this makes them bit-equal
        self.flow.u0.copy_(self.flow.u)
        self.contexts = [self.ctx]

    def exchange(self, field: torch.Tensor):
        """Ghost planes of a scalar or vector field from the neighbours (ifadv_exchange_planes; no-op on one GPU)."""
        if self.geom.world == 1:
            return
        ncomp = 1 if field.dim() == 3 else field.shape[-1]
        self.ctx.exchange_planes(torch.cuda.current_stream(field.device).cuda_stream, field.data_ptr(), ncomp, field.element_size())

    @property
    def bytes_sent(self) -> int:
        return self.ctx.slab_info()["bytes_sent"] if self.geom.world > 1 else 0

    @property
    def transport(self) -> str:
        return self.ctx.slab_info()["transport"] if self.geom.world > 1 else "none"

    def step(self, project=None):
        self.ia.mom_advect_step(self.flow, self.intf, 1.0, project=project)
        self.flow.dt.append(1.0)

    def mass(self) -> float:
        return self.ia.sum_inside(self.intf.f)  # a slab context sums its owned planes and all-reduces

    def owned_f(self) -> torch.Tensor:
        return self.intf.f[1:-1, 1:-1, self.geom.owned]

    def owned_rhou(self) -> torch.Tensor:
        return self.intf.rhou[1:-1, 1:-1, self.geom.owned, :]
