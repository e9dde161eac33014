"""Multi-GPU plumbing for the lattice MVM: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

The reference has no distributed code at all; this is new work defined by the
north star:

* ``broadcast_lattice`` -- the lattice (replay table, neighbour table, keys) is built once per
  hyper-parameter step on one rank and broadcast; every rank then filters its own RHS columns with no
  communication per MVM (columns are independent: splat, blur and slice are linear and per-channel).
* ``shard_columns`` -- which RHS columns a rank owns.
* ``PointShardedLattice`` -- for very large N: every rank holds the points ``[lo, hi)`` only.  The lattice is built
  SHARDED: each rank builds the lattice of its own points, the ranks' key lists are all-gathered and merged in rank
  order (``merge_key_lists``: that IS the reference's sequential first-touch numbering, because ranks own contiguous
  point ranges), and each rank keeps its own points' replay table in the global numbering.  splat is local, lattice
  values are combined with one all-reduce before the blur, slice is local.

Everything works on the ``gloo`` backend with CPU tensors for the host-logic tests (no GPU kernels are called
by the helpers that tests exercise on CPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_columns", "shard_points", "broadcast_lattice_arrays", "broadcast_lattice", "lattice_arrays",
           "all_gather_columns", "allreduce_lattice_values", "ColumnShardedOperator", "PointShardedLattice",
           "gather_key_lists", "merge_key_lists"]


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


_HEADER_WORDS = 8 + 16   # N, M, d, order, exact, n_coeffs, 2 spare, then up to 16 stencil coefficients as float32 bits


def _pack_header(meta: dict) -> torch.Tensor:
    import numpy as np
    c = np.asarray(meta["coeffs"], dtype=np.float32)
    if c.shape[0] > 16:
        raise ValueError("stencil too long for the broadcast header")
    h = torch.zeros(_HEADER_WORDS, dtype=torch.int64)
    h[:6] = torch.tensor([meta["N"], meta["M"], meta["d"], meta["order"], int(bool(meta.get("exact", False))), c.shape[0]])
    h[8:8 + c.shape[0]] = torch.from_numpy(c.view(np.int32).astype(np.int64))
    return h


def _unpack_header(h: torch.Tensor) -> dict:
    import numpy as np
    v = h.cpu().tolist()
    k = int(v[5])
    coeffs = np.asarray(v[8:8 + k], dtype=np.int64).astype(np.int32).view(np.float32).tolist()
    return {"N": int(v[0]), "M": int(v[1]), "d": int(v[2]), "order": int(v[3]), "exact": bool(v[4]), "coeffs": coeffs}


def _al(n: int) -> int:
    return (n + 255) & ~255


def broadcast_lattice_arrays(arrays: Optional[dict], meta: Optional[dict], src: int = 0, device=None, group=None):
    """Broadcast ``meta`` (N, M, d, order, coeffs, exact) and the lattice arrays from ``src``.

    On ``src`` pass the real ``arrays``/``meta``; elsewhere pass ``None``.  Returns ``(arrays, meta)`` on every rank.
    Two collectives: a 192-byte header (it sizes the receive buffer) and ONE byte buffer holding replay, keys and the
    neighbour table back to back -- three broadcasts plus a pickled ``broadcast_object_list`` cost 162 ms at 8 ranks in
    round 1 for 107 MB that need ~150 us of NVLink time.  As bytes: neither NCCL nor gloo has an int16 type, and a
    broadcast moves bits anyway."""
    rank = dist.get_rank(group)
    if rank == src:
        dev = arrays["replay"].device
        header = _pack_header(meta).to(dev)
    else:
        dev = torch.device(device) if device is not None else torch.device("cpu")
        header = torch.zeros(_HEADER_WORDS, dtype=torch.int64, device=dev)
    dist.broadcast(header, src=src, group=group)
    meta = _unpack_header(header)
    N, M, d, r = meta["N"], meta["M"], meta["d"], meta["order"]
    shapes = {"replay": (N, d + 1, 2), "keys": (M, d), "nbr": (d + 1, M, 2 * r)}
    sizes, offs, o = {}, {}, 0
    for name, dtype in _ARRAY_SPECS:
        n = 1
        for k in shapes[name]:
            n *= k
        sizes[name] = n * torch.empty((), dtype=dtype).element_size()
        offs[name] = o
        o += _al(sizes[name])
    buf = torch.empty(max(o, 1), dtype=torch.uint8, device=dev)
    if rank == src:
        for name, _ in _ARRAY_SPECS:
            if sizes[name]:
                buf[offs[name]:offs[name] + sizes[name]].copy_(arrays[name].contiguous().view(-1).view(torch.uint8))
    if o > 0:
        dist.broadcast(buf, src=src, group=group)
    out = {}
    for name, dtype in _ARRAY_SPECS:
        if rank == src:
            out[name] = arrays[name]
        elif sizes[name]:
            out[name] = buf[offs[name]:offs[name] + sizes[name]].view(dtype).view(shapes[name])
        else:
            out[name] = torch.empty(shapes[name], dtype=dtype, device=dev)
    return out, meta


def broadcast_lattice(lat, src: int = 0, device=None, group=None, build_csr: bool = False):
    """Broadcast a built ``Lattice`` from ``src``; other ranks pass ``lat=None`` and get a ``Lattice`` back (same
    structure arrays, same ``exact`` setting; the derived tables -- blur groups, row-sorted entries -- are rebuilt
    locally, which is cheaper than moving them)."""
    from .lattice import Lattice

    rank = dist.get_rank(group)
    if rank == src:
        meta = {"N": lat.N, "M": lat.M, "d": lat.d, "order": lat.order, "coeffs": lat.coeffs.tolist(), "exact": lat.exact}
        arrays = lattice_arrays(lat)
    else:
        meta, arrays = None, None
    arrays, meta = broadcast_lattice_arrays(arrays, meta, src=src, device=device, group=group)
    if rank == src:
        return lat
    return Lattice.from_arrays(meta["coeffs"], arrays["replay"], arrays["keys"], arrays["nbr"], build_csr=build_csr,
                               exact=meta["exact"])


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


def gather_key_lists(keys_local: torch.Tensor, group=None):
    """All-gather the ranks' key lists ``[M_g, d]`` int16 (ragged): returns the list of all ranks' key tensors, in rank
    order, identical on every rank.  One all-gather of the counts and one of the keys padded to the longest list (as
    bytes: no int16 collective type)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [keys_local]
    d = int(keys_local.shape[1])
    dev = keys_local.device
    count = torch.tensor([keys_local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    mmax = max(max(counts), 1)
    pad = torch.zeros((mmax, d), dtype=torch.int16, device=dev)
    pad[: keys_local.shape[0]] = keys_local
    parts = [torch.empty_like(pad).view(-1).view(torch.uint8) for _ in range(world)]
    dist.all_gather(parts, pad.view(-1).view(torch.uint8), group=group)
    return [p.view(torch.int16).view(mmax, d)[:c] for p, c in zip(parts, counts)]


def merge_key_lists(key_lists, want_map_of: Optional[int] = None):
    """Merge per-rank key lists (each in its rank's local first-touch order) into the global first-touch numbering:
    list 0, then the keys of list 1 that list 0 does not hold, ... (csrc: sgp_hash_append_keys / sgp_count_appended /
    sgp_number_appended).  Returns ``(keys[M, d], table, map)``: ``table`` is the key -> index hash table of the union
    (input of the neighbour / blur-group builders), ``map`` the int32 ``[M_g]`` local -> global index map of list
    ``want_map_of`` (or None).  CUDA tensors only -- there is no CPU path."""
    import ctypes as C

    from . import _capi
    from .lattice import _ptr, _stream_ptr
    lib = _capi.lib()
    first = key_lists[0]
    if not first.is_cuda:
        raise RuntimeError("merge_key_lists needs CUDA tensors: this package has no CPU path")
    dev, d = first.device, int(first.shape[1])
    total = sum(int(k.shape[0]) for k in key_lists)
    cap = int(lib.sgp_hash_capacity(max(total, 1)))
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        table = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        keys = torch.empty((max(total, 1), d), dtype=torch.int16, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        M, out_map = 0, None
        for g, kl in enumerate(key_lists):
            m = int(kl.shape[0])
            if m == 0:
                if want_map_of == g:
                    out_map = torch.empty(0, dtype=torch.int32, device=dev)
                continue
            kl = kl.contiguous()
            slot_of = torch.empty(m, dtype=torch.int32, device=dev)
            ws_bytes = int(lib.sgp_number_workspace_bytes(m, 0))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _capi.check(lib.sgp_hash_append_keys(_ptr(kl), m, d, _ptr(keys), M, _ptr(table), cap, _ptr(slot_of), _ptr(flags),
                                                 st))
            m_add, fl = C.c_int64(0), C.c_int32(0)
            _capi.check(lib.sgp_count_appended(_ptr(table), cap, _ptr(slot_of), m, _ptr(ws), ws_bytes, _ptr(flags),
                                               C.byref(m_add), C.byref(fl), st))
            mp = torch.empty(m, dtype=torch.int32, device=dev) if want_map_of == g else None
            _capi.check(lib.sgp_number_appended(_ptr(table), cap, _ptr(slot_of), _ptr(kl), m, d, _ptr(ws), M, int(m_add.value),
                                                _ptr(mp), _ptr(keys), st))
            if mp is not None:
                out_map = mp
            M += int(m_add.value)
    return keys[:M], table, out_map


class PointShardedLattice:
    """Point sharding for very large N: rank ``k`` owns the points ``[lo_k, hi_k)`` and builds ONLY their part of the
    lattice (per-point stage, hash insertion, replay table: everything sized by ``N (d+1)`` is sharded).

    Build: (1) the lattice of the rank's own points -> local keys in local first-touch order; (2) all-gather of the key
    lists (``M_g * d * 2`` bytes each); (3) every rank merges the lists in rank order, which reproduces the reference's
    sequential first-touch numbering of the whole point set bit for bit (``merge_key_lists``); (4) the rank's replay
    table is re-indexed to the global numbering and the tables every rank needs whole (blur groups / neighbour table,
    built from the merged key table) are built locally.  One MVM is: local splat of the rank's rows of ``V`` into the
    full ``[M, L]`` lattice, **one all-reduce (sum) of ``M*L*4`` bytes**, blur (replicated), local slice of the rank's
    rows.  ``mvm`` takes and returns this rank's row block ``[hi-lo, L]``.

    ``x``: the rank's own points ``[hi-lo, d]`` (``x_is_local=True``) or the whole point set, of which the rank takes
    its share (contiguous blocks, ``shard_points``)."""

    def __init__(self, x: torch.Tensor, coeffs, group=None, x_is_local: bool = False, **lattice_kwargs):
        from .lattice import Lattice

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if x_is_local:
            n_loc = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
            if self.world > 1:
                counts = [torch.zeros_like(n_loc) for _ in range(self.world)]
                dist.all_gather(counts, n_loc, group=group)
                counts = [int(c.item()) for c in counts]
            else:
                counts = [int(x.shape[0])]
            self.lo = sum(counts[: self.rank])
            self.hi = self.lo + counts[self.rank]
            self.N = sum(counts)
            x_loc = x
        else:
            self.N = int(x.shape[0])
            self.lo, self.hi = shard_points(self.N, self.world, self.rank)
            x_loc = x[self.lo:self.hi]
        build_nbr = lattice_kwargs.pop("build_nbr", True)
        mine = Lattice(x_loc.contiguous(), coeffs, build_groups=False, build_rows=False, build_nbr=False,
                       keep_structure=False)
        self.d, self.M_local = mine.d, mine.M
        lists = gather_key_lists(mine.keys, group=group)
        keys, table, lmap = merge_key_lists(lists, want_map_of=self.rank)
        self.M = int(keys.shape[0])
        replay = mine.replay
        if replay.numel():
            replay[..., 0] = lmap[replay[..., 0].long()]      # local -> global lattice indices
        del lists, lmap
        self.local = Lattice.from_arrays(mine.coeffs, replay, keys.contiguous(), None, table=table, build_nbr=build_nbr,
                                         **lattice_kwargs)
        del mine, table

    def mvm(self, V_local: torch.Tensor, out: Optional[torch.Tensor] = None, column_blur: Optional[bool] = None,
            **kw) -> torch.Tensor:
        """This rank's rows of the product.  ``column_blur=True`` (needs ``L % world == 0`` and the blur groups) replaces
        the all-reduce + replicated blur by reduce-scatter over COLUMN blocks -> blur of this rank's ``L / world`` columns
        -> all-gather: the same bytes on the wire, and the blur -- the part of the step that point sharding alone leaves
        replicated -- is divided by the world size.  Off by default: measured SLOWER on 2 B200 at the stress shape D/10
        (L = 4: 8.2 ms against 6.2 ms per MVM) -- the two re-layout copies move the lattice values in 8-byte pieces and
        the blur-group kernel is row-count-bound, so two columns cost it almost what four do."""
        if V_local.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns {self.hi - self.lo} points, got {V_local.shape[0]} rows")
        L = int(V_local.shape[1])
        lat = self.local
        can = (self.world > 1 and L % self.world == 0 and lat.groups is not None and lat.rows is not None and not kw
               and V_local.dtype == torch.float32)
        if column_blur is None:
            column_blur = False
        if column_blur and not can:
            raise ValueError("column_blur needs world > 1, L % world == 0, blur groups and row-sorted entries")
        if column_blur:
            return self._mvm_column_blur(V_local, out)
        hook = (lambda vals: allreduce_lattice_values(vals, group=self.group)) if self.world > 1 else None
        return lat.mvm(V_local, out=out, after_splat=hook, **kw)

    def _mvm_column_blur(self, V: torch.Tensor, out: Optional[torch.Tensor]) -> torch.Tensor:
        import ctypes as C

        from . import _capi
        from .lattice import _fp, _ptr, _stream_ptr
        lib, lat, G = _capi.lib(), self.local, self.world
        dev, M, L = lat.device, lat.M, int(V.shape[1])
        Lc = L // G
        n_loc = int(V.shape[0])
        if V.stride(1) != 1 and L > 1:
            V = V.contiguous()
        if out is None:
            out = torch.empty((n_loc, L), dtype=torch.float32, device=dev)
        key = ("colblur", L)
        bufs = lat._bufs.get(key)
        if bufs is None:
            bufs = {"full": torch.empty((M, L), dtype=torch.float32, device=dev),
                    "parts": torch.empty((G, M, Lc), dtype=torch.float32, device=dev),
                    "b0": torch.empty((M, Lc), dtype=torch.float32, device=dev),
                    "b1": torch.empty((M, Lc), dtype=torch.float32, device=dev)}
            lat._bufs = {key: bufs}
        full, parts, b0, b1 = bufs["full"], bufs["parts"], bufs["b0"], bufs["b1"]
        c = lat.coeffs
        where = C.c_int(0)
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            if n_loc > 0:
                _capi.check(lib.sgp_splat_rows(_ptr(lat.rows["ent"]), _ptr(lat.rows["seg_row"]), lat.rows["n"], lat.N, M, _ptr(V),
                                               V.stride(0), L, _ptr(full), L, st))
            else:
                full.zero_()
            parts.copy_(full.view(M, G, Lc).permute(1, 0, 2))               # column blocks made contiguous
            dist.reduce_scatter_tensor(b0, parts, group=self.group)         # this rank's columns, summed over the ranks
            arr = lat.groups["array"]
            _capi.check(lib.sgp_blur_groups(arr, len(arr), M, lat.order, _fp(c), c.shape[0], Lc, _ptr(b0), _ptr(b1),
                                            C.byref(where), 0 if lat.exact else 1, st))
            dist.all_gather_into_tensor(parts, b1 if where.value else b0, group=self.group)
            full.view(M, G, Lc).copy_(parts.permute(1, 0, 2))               # back to [M, L] rows, last stage's order
            if n_loc > 0:
                v_out = lat._view(lat._table(False, True), None, lat.exact)
                _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(full), L, _ptr(out), out.stride(0), L, st))
        return out
