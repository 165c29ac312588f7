"""z-slab decomposition of the VOF + CMOM advection path across the GPUs of one box (one process per GPU).

The reference has no multi-device code; this is new functionality whose oracle is the single-GPU
run on the same global grid: owned cells must come out BIT-IDENTICAL.

Scheme (communication-avoiding wide halo): every rank runs the UNCHANGED single-GPU C ABI on its slab extended by
W overlap planes per interior side and treats the slab ends as ordinary (non-periodic) boundaries.  The wrong
boundary rule there contaminates only overlap planes: one directional sweep along z moves the error 3 planes up /
2 planes down, a sweep along x or y 2 up / 1 down (stencil reach of f2face!/ϕu/PLIC), so one
advectfq! (3 sweeps) reaches at most 7 planes from the lower and 4 from the upper slab end.  Both advectfq! calls of
a step start from exchanged fields (ρu is rebuilt from u and f by u2ρu! each time), so ONE exchange of W = 8 planes of
f per step (plus u when the caller's projection changed it) keeps every owned cell exact.  The exchange is 2·W
contiguous planes per neighbour (z is the slowest index) posted as NCCL send/recv over NVLink; at 512² planes that is
8.5 MB against >100 ms of sweep compute per step, so it is not worth splitting the sweep to overlap it.
The price is the redundant update of 2·W overlap planes (3 % at 512 owned planes per GPU).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

W_DEFAULT = 8  # >= 7 (lower) / 4 (upper) planes one advectfq! can contaminate, + 1 for the stale ghost plane


@dataclass
class SlabGeom:
    """Geometry of rank `rank`'s slab: the global grid is N[0] x N[1] x (nz_owned * world)."""
    rank: int
    world: int
    nz_owned: int
    W: int
    per_z: bool

    @property
    def wlo(self) -> int:
        return self.W if (self.world > 1 and (self.rank > 0 or self.per_z)) else 0

    @property
    def whi(self) -> int:
        return self.W if (self.world > 1 and (self.rank < self.world - 1 or self.per_z)) else 0

    @property
    def nz_local(self) -> int:  # interior planes of the local array (owned + overlap)
        return self.nz_owned + self.wlo + self.whi

    @property
    def z_origin(self) -> int:  # global 0-based index of the first local interior plane
        return self.rank * self.nz_owned - self.wlo

    @property
    def owned(self) -> slice:  # owned planes as a slice of the local array's z axis (ghost plane at 0)
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
    """Contiguous view of planes z0:z1 of a column-major (x,y,z[,c]) field -- scalar fields only give one block."""
    assert t.dim() == 3
    return t.permute(2, 1, 0)[z0:z1]  # storage order is (z,y,x): a z-range is one contiguous block


def exchange_overlap(fields: List[torch.Tensor], g: SlabGeom, dist=None) -> int:
    """Fill the overlap planes of every scalar field from the neighbours' owned planes.  Returns bytes sent.

    Upward message: my top W owned planes -> the upper neighbour's lower overlap.  Downward: my bottom W owned planes ->
    the lower neighbour's upper overlap.  All sends/recvs of one call go out as ONE batch (ncclGroupStart/End)."""
    if g.world == 1:
        return 0
    import torch.distributed as tdist

    dist = dist or tdist
    sends, recvs, sent = [], [], 0
    o = g.owned
    # Posting order matters when both neighbours are the same peer (2 ranks, periodic z): NCCL matches the k-th send to a
    # peer with that peer's k-th recv, so sends go [up, down] and recvs [from below (an up-message), from above].
    for f in fields:
        if g.upper is not None:
            top = zplanes(f, o.stop - g.W, o.stop)
            sends.append(dist.P2POp(dist.isend, top, g.upper))
            sent += top.numel() * top.element_size()
        if g.lower is not None:
            bot = zplanes(f, o.start, o.start + g.W)
            sends.append(dist.P2POp(dist.isend, bot, g.lower))
            sent += bot.numel() * bot.element_size()
        if g.lower is not None and g.wlo:
            recvs.append(dist.P2POp(dist.irecv, zplanes(f, o.start - g.wlo, o.start), g.lower))
        if g.upper is not None and g.whi:
            recvs.append(dist.P2POp(dist.irecv, zplanes(f, o.stop, o.stop + g.whi), g.upper))
    ops = sends + recvs
    if not ops:
        return 0
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return sent


class SlabRunner:
    """One rank of the slab-decomposed CMOM advection benchmark / simulation (CUDA)."""

    def __init__(self, N_per_gpu, dtype: str, perdir, kind: str, rank: int, world: int, device, W: int = W_DEFAULT, lam_rho=None,
                 fields=None):
        """fields: optional (f_local, u_local) column-major device tensors (e.g. slices of a global state); by default the
        slab samples the analytic generators of `kind` at its global offset."""
        import interfaceadvection.jl_b200 as ia
        from . import configs

        self.ia = ia
        N1, N2, nz = N_per_gpu
        self.geom = g = SlabGeom(rank, world, nz, W, 3 in tuple(perdir))
        self.Nglobal = (N1, N2, nz * world)
        self.Nlocal = (N1, N2, g.nz_local)
        self.perdir = g.local_perdir(perdir)
        T = getattr(torch, dtype)
        origin = (0, 0, g.z_origin)
        self.flow = ia.Flow(self.Nlocal, (0, 0, 0), T=T, dt=1.0, perdir=self.perdir, device=device)
        if fields is None:
            case = configs.make_case(self.Nglobal, dtype=dtype, device=device, kind=kind, Nl=self.Nlocal, origin=origin)
            self.lam_rho = case["lam_rho"] if lam_rho is None else lam_rho
            self.intf = ia.cVOF(self.Nlocal, T=T, InterfaceSDF=case["sdf"], lam_rho=self.lam_rho, perdir=self.perdir, device=device,
                                origin=origin)
            self.flow.u.copy_(case["u"])
        else:
            self.lam_rho = 1e-3 if lam_rho is None else lam_rho
            self.intf = ia.cVOF(self.Nlocal, T=T, lam_rho=self.lam_rho, perdir=self.perdir, device=device)
            self.intf.f.copy_(fields[0])
            ia.BCf(self.intf.f, self.perdir)
            self.flow.u.copy_(fields[1])
        ia.BC(self.flow.u, (0, 0, 0), False, self.perdir)
        self.flow.u0.copy_(self.flow.u)
        self.contexts = [ia.context_for(self.intf.f)]
        self.bytes_sent = 0
        self.exchange()  # overlap planes of the initial f come from the analytic SDF already; this makes them bit-equal

    def exchange(self):
        self.bytes_sent += exchange_overlap([self.intf.f], self.geom)

    def step(self):
        self.ia.mom_advect_step(self.flow, self.intf, 1.0)
        self.flow.dt.append(1.0)
        self.exchange()

    def mass(self) -> float:
        o = self.geom.owned
        m = self.intf.f[1:-1, 1:-1, o].sum(dtype=torch.float64)
        if self.geom.world > 1:
            import torch.distributed as dist

            dist.all_reduce(m, op=dist.ReduceOp.SUM)
        return float(m.item())

    def owned_f(self) -> torch.Tensor:
        return self.intf.f[1:-1, 1:-1, self.geom.owned]

    def owned_rhou(self) -> torch.Tensor:
        return self.intf.rhou[1:-1, 1:-1, self.geom.owned, :]
