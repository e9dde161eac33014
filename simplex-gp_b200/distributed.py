"""Multi-GPU plumbing for the lattice MVM: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

The reference has no distributed code at all; this is new work defined by the
north star:

* ``broadcast_lattice`` -- the lattice (replay table, neighbour table, keys) is built once per
  hyper-parameter step on one rank and broadcast; every rank then filters its own RHS columns with no
  communication per MVM (columns are independent: splat, blur and slice are linear and per-channel).
* ``shard_columns`` -- which RHS columns a rank owns.
* ``PointShardedLattice`` -- for very large N: every rank holds the points ``[lo, hi)`` and the full lattice
  numbering; splat is local, lattice values are combined with one all-reduce before the blur, slice is local.

Everything works on the ``gloo`` backend with CPU tensors for the host-logic tests (no GPU kernels are called
by the helpers that tests exercise on CPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_columns", "shard_points", "broadcast_lattice_arrays", "broadcast_lattice", "lattice_arrays",
           "all_gather_columns", "allreduce_lattice_values", "ColumnShardedOperator", "PointShardedLattice"]


def shard_columns(L: int, world: int, rank: int) -> Tuple[int, int]:
    """Half-open column range ``[lo, hi)`` of rank ``rank`` when ``L`` RHS columns are split over ``world`` ranks
    (earlier ranks take the remainder; ranks beyond ``L`` get an empty range)."""
    if L < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard request L={L} world={world} rank={rank}")
    base, extra = divmod(L, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_points(N: int, world: int, rank: int) -> Tuple[int, int]:
    """Half-open point range of a rank for point sharding (contiguous blocks, same rule as ``shard_columns``)."""
    return shard_columns(N, world, rank)


_ARRAY_SPECS = (
    # name, dtype
    ("replay", torch.int32),
    ("keys", torch.int16),
    ("nbr", torch.int32),
)


def lattice_arrays(lat) -> dict:
    """The arrays that define a built lattice for the MVM (what has to travel to the other ranks)."""
    if lat.nbr is None:
        raise RuntimeError("a lattice built with build_nbr=False cannot be broadcast: the receiving ranks rebuild "
                           "their blur groups from the neighbour table")
    return {"replay": lat.replay, "keys": lat.keys, "nbr": lat.nbr}


def broadcast_lattice_arrays(arrays: Optional[dict], meta: Optional[dict], src: int = 0, device=None, group=None):
    """Broadcast ``meta`` (small dict: N, M, d, order, coeffs) and the lattice arrays from ``src``.

    On ``src`` pass the real ``arrays``/``meta``; elsewhere pass ``None``.  Returns ``(arrays, meta)`` on every
    rank.  One ``broadcast_object_list`` for the metadata and one ``broadcast`` per array."""
    rank = dist.get_rank(group)
    box = [meta if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    meta = box[0]
    N, M, d, r = meta["N"], meta["M"], meta["d"], meta["order"]
    shapes = {"replay": (N, d + 1, 2), "keys": (M, d), "nbr": (d + 1, M, 2 * r)}
    out = {}
    for name, dtype in _ARRAY_SPECS:
        if rank == src:
            t = arrays[name].contiguous()
        else:
            t = torch.empty(shapes[name], dtype=dtype, device=device)
        if t.numel() > 0:
            # as bytes: neither NCCL nor gloo has an int16 type, and a broadcast moves bits anyway
            dist.broadcast(t.view(-1).view(torch.uint8), src=src, group=group)
        out[name] = t
    return out, meta


def broadcast_lattice(lat, src: int = 0, device=None, group=None, build_csr: bool = False):
    """Broadcast a built ``Lattice`` from ``src``; other ranks pass ``lat=None`` and get a ``Lattice`` back."""
    from .lattice import Lattice

    rank = dist.get_rank(group)
    if rank == src:
        meta = {"N": lat.N, "M": lat.M, "d": lat.d, "order": lat.order, "coeffs": lat.coeffs.tolist()}
        arrays = lattice_arrays(lat)
    else:
        meta, arrays = None, None
    arrays, meta = broadcast_lattice_arrays(arrays, meta, src=src, device=device, group=group)
    if rank == src:
        return lat
    return Lattice.from_arrays(meta["coeffs"], arrays["replay"], arrays["keys"], arrays["nbr"], build_csr=build_csr)


def all_gather_columns(mine: torch.Tensor, L: int, group=None) -> torch.Tensor:
    """Reassemble ``[N, L]`` from per-rank column blocks laid out by ``shard_columns`` (ragged blocks allowed).

    Column-sharded MVMs need this only when the caller wants the full block on every rank; CG keeps its RHS
    column-sharded and all-reduces ``L`` dot products instead."""
    world = dist.get_world_size(group)
    N = mine.shape[0]
    widths = [hi - lo for lo, hi in (shard_columns(L, world, r) for r in range(world))]
    wmax = max(widths) if widths else 0
    if wmax == 0:
        return mine.new_empty((N, 0))
    pad = mine.new_zeros((N, wmax))
    pad[:, : mine.shape[1]] = mine
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:, :w] for p, w in zip(parts, widths)], dim=1)


def allreduce_lattice_values(values: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank splatted lattice values ``[M, L]`` in place (point sharding: the one exchange step of the
    MVM, between splat and blur; ``M*L*4`` bytes over NVLink with the NCCL backend)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
    return values


class ColumnShardedOperator:
    """``K @ V`` with the RHS columns of ``V[N, L]`` split over the ranks (the default multi-GPU form of the north
    star).  Every rank holds the same lattice (built on ``src`` and broadcast, or built locally from the same ``x``)
    and filters its own columns; there is no communication inside an MVM.  ``matmul`` returns this rank's column block;
    ``matmul_full`` all-gathers the blocks (only needed when the caller wants the whole product everywhere -- a CG
    solver keeps its vectors column-sharded and all-reduces its dot products instead, see ``dots``)."""

    def __init__(self, lat, group=None):
        self.lat = lat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def columns(self, L: int) -> Tuple[int, int]:
        return shard_columns(L, self.world, self.rank)

    def matmul(self, V_local: torch.Tensor) -> torch.Tensor:
        if V_local.shape[1] == 0:
            return V_local.new_empty(V_local.shape)
        return self.lat.mvm(V_local.contiguous())

    def matmul_full(self, V: torch.Tensor) -> torch.Tensor:
        lo, hi = self.columns(V.shape[1])
        mine = self.matmul(V[:, lo:hi])
        if self.world == 1:
            return mine
        return all_gather_columns(mine, V.shape[1], group=self.group)

    def dots(self, A_local: torch.Tensor, B_local: torch.Tensor, L: int) -> torch.Tensor:
        """Column-wise dot products ``[L]`` of two column-sharded blocks, available on every rank (one all-reduce of
        ``L`` floats: the only collective a column-sharded CG iteration needs)."""
        lo, hi = self.columns(L)
        out = A_local.new_zeros(L)
        out[lo:hi] = (A_local * B_local).sum(0)
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out


class PointShardedLattice:
    """Point sharding for very large N: rank ``k`` owns the points ``[lo_k, hi_k)``.

    The lattice numbering has to be global, so every rank runs the (deterministic) lattice build on all ``N`` points
    -- positions are ``N*d*4`` bytes, small next to the value traffic -- and then keeps the per-point tables (replay,
    row-sorted entries) of its own points only.  One MVM is: local splat of the rank's rows of ``V`` into the full
    ``[M, L]`` lattice, **one all-reduce (sum) of ``M*L*4`` bytes**, blur (replicated), local slice of the rank's rows.
    ``mvm`` takes and returns this rank's row block ``[hi-lo, L]``."""

    def __init__(self, x_full: torch.Tensor, coeffs, group=None, **lattice_kwargs):
        from .lattice import Lattice

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        N = int(x_full.shape[0])
        self.lo, self.hi = shard_points(N, self.world, self.rank)
        full = Lattice(x_full, coeffs, build_groups=False, build_rows=False, sort_points=False, build_tiles=False)
        self.N, self.M, self.d = full.N, full.M, full.d
        # the same lattice restricted to this rank's points: shares keys and the neighbour table, rebuilds the
        # per-point tables for hi-lo points
        self.local = Lattice.from_arrays(full.coeffs, full.replay[self.lo:self.hi].contiguous(), full.keys, full.nbr,
                                         **lattice_kwargs)
        del full

    def mvm(self, V_local: torch.Tensor, **kw) -> torch.Tensor:
        if V_local.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns {self.hi - self.lo} points, got {V_local.shape[0]} rows")
        hook = (lambda vals: allreduce_lattice_values(vals, group=self.group)) if self.world > 1 else None
        return self.local.mvm(V_local, after_splat=hook, **kw)
