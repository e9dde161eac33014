"""Host side of the lattice filter: a ``Lattice`` built once per ``(x / lengthscale, stencil)`` and
reused by every MVM, plus ``lattice_filter`` -- the one-call form with the reference's signature.

Reference being replaced: ``PermutohedralLattice::filter`` (gpytorch_lattice_kernel/cpp/permutohedral.h:259-340),
reached through ``filter(src, ref, coeffs)`` (cpp/lattice.cpp:6-16, bilateral_kernel.py:95).  The reference
rebuilds the whole lattice inside every call; here the structure (``replay``, ``keys``, ``nbr``) lives in HBM as
flat torch tensors and an MVM is splat -> (d+1) x blur -> slice over two ``[M, L]`` ping-pong buffers.

PyTorch is used for device memory and streams only; all compute goes through the C ABI in ``_capi``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi
from ._capi import LatticeView, TilesView, check

__all__ = ["Lattice", "lattice_filter", "stencil_variance", "scale_factors", "slice_divisor"]


def _coeffs_np(coeffs) -> np.ndarray:
    if isinstance(coeffs, torch.Tensor):
        coeffs = coeffs.detach().to("cpu", torch.float32).numpy()
    c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float32))
    if c.ndim != 1 or c.shape[0] % 2 != 1:
        raise ValueError(f"stencil coefficients must be a 1-D array of odd length, got shape {c.shape}")
    if c.shape[0] // 2 > _capi.SGP_MAX_ORDER:
        raise ValueError(f"stencil order {c.shape[0] // 2} > {_capi.SGP_MAX_ORDER}")
    return c


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def stencil_variance(coeffs) -> np.float32:
    """Second central moment of the stencil in index units (permutohedral.h:203-219), fp32."""
    c = _coeffs_np(coeffs)
    out = C.c_float(0.0)
    check(_capi.lib().sgp_stencil_variance(_fp(c), c.shape[0], C.byref(out)))
    return np.float32(out.value)


def scale_factors(d: int, var) -> np.ndarray:
    """Per-axis scale of the elevation (permutohedral.h:372-390), fp32 ``[d]``."""
    out = np.empty(d, dtype=np.float32)
    check(_capi.lib().sgp_scale_factors(int(d), C.c_float(float(var)), _fp(out)))
    return out


def slice_divisor(d: int) -> np.float32:
    return np.float32(_capi.lib().sgp_slice_divisor(int(d)))


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


_GROUP_HINTS = {}   # (d, order, rows per CTA) -> axis-range lengths of the last blur-group build of that shape
_M_BUILD_HINT = {}  # (N, d) -> lattice points of the last build of that shape: sizes the next build's hash table


class Lattice:
    """Permutohedral lattice of the points ``x[N, d]`` (already divided by the lengthscale).

    Device arrays (all torch tensors on ``x.device``), in the reference's numbering -- lattice point ``i`` is
    the ``i``-th key the reference's sequential ``splat`` loop would have created:

    ========  ==================  =========================================================================
    greedy    int16 ``[N, d+1]``   remainder-0 lattice point of each input point (permutohedral.h:404-457)
    rank      int8  ``[N, d+1]``   simplex permutation (:425-457)
    replay    int32 ``[N, d+1, 2]`` ``{lattice index, fp32 weight bits}`` per simplex vertex (:482-483)
    keys      int16 ``[M, d]``     lattice keys in first-touch order (:73-79, :470-471)
    nbr       int32 ``[d+1, M, 2r]`` blur neighbours, ``t`` over ``o = -r..-1, 1..r``; -1 = absent (:541-545)
    ========  ==================  =========================================================================

    Derived tables for the MVM kernels (internal orderings; none of them changes the numbering above):

    * ``rows``    point-vertices sorted by lattice row, ``ent[9M', 2] {point | row-start flag, weight}`` / ``seg_row[9M'/4]``: the
                  segmented-gather splat (``build_rows``, default on); ``csr_ptr`` adds row starts for the ordered gather
                  of ``mode=2`` (``build_csr``).
    * ``groups``  blur groups: for each range of consecutive axes, lattice points sorted by class, CTA batches, gather
                  list and batch-local neighbour table (``build_groups``, default on; ``group_axes=None`` sizes the
                  ranges adaptively, ``group_rows`` is the CTA capacity).
    * ``sorted`` / ``tiles``  locality order of the points and shared-memory tiles (``sort_points`` / ``build_tiles``,
                  default off: measured slower on B200, kept for comparison).

    ``exact=True`` makes ``mvm`` default to the reference's arithmetic (see ``mvm``); ``keep_structure=False`` drops
    ``greedy`` / ``rank`` after the build; ``hash_capacity`` overrides the table size (tests).
    """

    def __init__(self, x: torch.Tensor, coeffs, *, build_csr: bool = False, build_tiles: bool = False,
                 build_groups: bool = True, group_axes: Optional[int] = None, group_rows: int = 512,
                 sort_points: bool = False, build_rows: bool = True, build_nbr: bool = True, exact: bool = False,
                 tile_points: int = 256, keep_structure: bool = True, hash_capacity: Optional[int] = None):
        if x.dim() != 2:
            raise ValueError(f"x must be [N, d], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("Lattice needs a CUDA tensor: this package has no CPU path")
        if x.dtype != torch.float32:
            raise TypeError(f"x must be float32 (the reference CPU filter is fp32-only), got {x.dtype}")
        lib = _capi.lib()
        x = x.detach().contiguous()
        self.device = x.device
        self.N, self.d = int(x.shape[0]), int(x.shape[1])
        if not (1 <= self.d <= _capi.SGP_MAX_DIM):
            raise ValueError(f"d={self.d} outside [1, {_capi.SGP_MAX_DIM}]")
        self.coeffs = _coeffs_np(coeffs)
        self.order = self.coeffs.shape[0] // 2
        self.var = stencil_variance(self.coeffs)
        self.scale = scale_factors(self.d, self.var)
        self.exact = bool(exact)
        N, d, r = self.N, self.d, self.order
        dev = self.device
        total = N * (d + 1)
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            self.greedy = torch.empty((N, d + 1), dtype=torch.int16, device=dev)
            self.rank = torch.empty((N, d + 1), dtype=torch.int8, device=dev)
            self.replay = torch.empty((N, d + 1, 2), dtype=torch.int32, device=dev)
            flags = torch.zeros(1, dtype=torch.int32, device=dev)
            self.M = 0
            self.keys = torch.empty((0, d), dtype=torch.int16, device=dev)
            self.nbr = torch.empty((d + 1, 0, 2 * r), dtype=torch.int32, device=dev)
            self.csr_ptr = None
            self.csr_ent = None
            self.tiles = None
            self.groups = None
            self.sorted = None
            self.rows = None
            self.hash_capacity = 0
            if N > 0:
                check(lib.sgp_build_points(_ptr(x), N, d, x.stride(0), _fp(self.scale), _ptr(self.greedy),
                                           _ptr(self.rank), _ptr(self.replay), _ptr(flags), st))
                # The table only has to hold the DISTINCT keys (M of them, 4 % of the N(d+1) insertions at the metric
                # shape), but M is not known before the insertion.  Successive hyper-parameter steps build almost the
                # same lattice, so the table is sized for four times the M of the last build of this shape (8 MB instead
                # of 268 MB at the metric shape: it then lives in L2 for the 9 M atomic probes, and the fill / re-target
                # passes over it shrink with it); if the lattice outgrew that, the insertion reports a full table and is
                # repeated at the safe size 2 N(d+1).  The numbering does not depend on the table's size or layout.
                cap_full = int(lib.sgp_hash_capacity(total))
                hint = _M_BUILD_HINT.get((N, d))
                cap = int(hash_capacity) if hash_capacity else (
                    min(cap_full, int(lib.sgp_hash_capacity(2 * hint))) if hint else cap_full)
                slot_of = torch.empty(total, dtype=torch.int32, device=dev)
                ws_bytes = int(lib.sgp_number_workspace_bytes(N, d))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                M = C.c_int64(0)
                fl = C.c_int32(0)
                while True:
                    table = torch.full((cap,), -1, dtype=torch.int64, device=dev)
                    check(lib.sgp_hash_insert(_ptr(self.greedy), _ptr(self.rank), N, d, _ptr(table), cap, _ptr(slot_of),
                                              _ptr(flags), st))
                    rc = lib.sgp_count_points(_ptr(table), cap, _ptr(slot_of), N, d, _ptr(ws), ws_bytes, _ptr(flags),
                                              C.byref(M), C.byref(fl), st)
                    if rc == -3 and (fl.value & 2) and cap < cap_full and not hash_capacity:   # SGP_FLAG_TABLE_FULL
                        cap = cap_full
                        flags.zero_()
                        del table
                        continue
                    check(rc)
                    if 2 * int(M.value) > cap and cap < cap_full and not hash_capacity:
                        cap = cap_full       # over half full: the neighbour look-ups that follow would probe long chains
                        flags.zero_()
                        del table
                        continue
                    break
                self.hash_capacity = cap
                self.M = int(M.value)
                _M_BUILD_HINT[(N, d)] = self.M
                self._covers_all_rows = True      # every lattice point was created by one of these points
                self.keys = torch.empty((self.M, d), dtype=torch.int16, device=dev)
                check(lib.sgp_number_points(_ptr(table), cap, _ptr(slot_of), _ptr(self.greedy), _ptr(self.rank), N, d,
                                            _ptr(ws), self.M, _ptr(self.replay), _ptr(self.keys), st))
                del slot_of, ws
                self._build_derived(table, build_nbr=build_nbr, build_csr=build_csr, build_groups=build_groups,
                                    group_axes=group_axes, group_rows=group_rows, sort_points=sort_points,
                                    build_tiles=build_tiles, tile_points=tile_points, build_rows=build_rows)
                del table
            if not keep_structure:
                self.greedy = None
                self.rank = None
        self._bufs = {}
        self._tables = {}

    def _build_derived(self, table, *, build_nbr=True, build_csr=False, build_groups=True, group_axes=None,
                       group_rows=512, sort_points=False, build_tiles=False, tile_points=256, build_rows=True) -> None:
        """Everything an MVM reads beyond ``replay``: neighbour table, blur groups, row-sorted entries (and the optional
        CSR / locality order / tiles).  ``table``: the key -> lattice index hash table left by the numbering."""
        lib = _capi.lib()
        dev, d, r = self.device, self.d, self.order
        st = _stream_ptr(dev)
        cap = int(table.numel())
        if build_nbr:
            self.nbr = torch.empty((d + 1, self.M, 2 * r), dtype=torch.int32, device=dev)
            check(lib.sgp_build_neighbours(_ptr(self.keys), self.M, d, r, _ptr(table), cap, _ptr(self.nbr), st))
        else:
            self.nbr = None    # the blur groups are built straight from the hash table
        if build_csr:
            self._build_csr()
        pending = None
        if build_groups and r >= 1 and self.M > 0:
            # the axis ranges of the last lattice of this shape: every group launched back to back, no host round trip,
            # the batch packing on a second stream beside the next group's class sort and the row sort below
            pending = self._launch_groups_hinted(group_axes, group_rows, table=None if build_nbr else table)
            if pending is not None and build_rows and self.rows is None and not (sort_points or build_tiles or build_csr):
                self._build_rows()
            if pending is None or not self._finish_groups_hinted(pending):
                self._build_groups(group_axes, group_rows, table=None if build_nbr else table)
        if not build_nbr and build_groups and r >= 1 and self.groups is None:   # no groups (long 1-D lines): the per-axis blur needs nbr
            self.nbr = torch.empty((d + 1, self.M, 2 * r), dtype=torch.int32, device=dev)
            check(lib.sgp_build_neighbours(_ptr(self.keys), self.M, d, r, _ptr(table), cap, _ptr(self.nbr), st))
        if (sort_points or build_tiles) and self.M > 0 and self.greedy is not None:
            self._sort_points()
        if build_tiles and self.M > 0 and self.sorted is not None:
            self._build_tiles(tile_points)
        if build_rows and self.M > 0 and self.rows is None:
            self._build_rows()

    def extend(self, x_new: torch.Tensor, *, build_groups: bool = True, group_axes: Optional[int] = None,
               group_rows: int = 512, build_rows: bool = True, build_nbr: bool = True,
               build_csr: bool = False, lazy_tables: int = 0) -> "Lattice":
        """The lattice of ``cat([x, x_new])`` without revisiting ``x``: a new ``Lattice`` whose first ``N`` points and
        first ``M`` lattice points are this lattice's, bit for bit what ``Lattice(torch.cat([x, x_new]), coeffs)``
        builds (first-touch numbering is sequential over the points).  This lattice is not modified.

        It is the union lattice of the reference's rectangular operator (bilateral_kernel.py:142-160) with the training
        lattice reused: the per-point stage, the hash insertion and the numbering run over ``x_new`` only; the tables
        that depend on the whole key set (neighbours, blur groups, row-sorted entries) are rebuilt.

        ``lazy_tables=k > 0`` postpones the blur groups and the row-sorted entries (2.0 of the 2.3 ms of an extension at
        the metric shape) until the ``k+1``-th product: the first ``k`` run on the neighbour table alone (atomic splat,
        one blur launch per axis -- about 0.5 ms instead of 0.2 ms each), which is the cheaper total for the one or
        two products of a prediction."""
        if x_new.dim() != 2 or int(x_new.shape[1]) != self.d:
            raise ValueError(f"x_new must be [N_new, {self.d}], got {tuple(x_new.shape)}")
        if not x_new.is_cuda or x_new.device != self.device:
            raise RuntimeError(f"x_new must live on {self.device}")
        if x_new.dtype != torch.float32:
            raise TypeError(f"x_new must be float32, got {x_new.dtype}")
        lib = _capi.lib()
        dev, d, r = self.device, self.d, self.order
        x_new = x_new.detach().contiguous()
        Nn = int(x_new.shape[0])
        new = object.__new__(type(self))
        new.device, new.d, new.order = dev, d, r
        new.coeffs, new.var, new.scale, new.exact = self.coeffs, self.var, self.scale, self.exact
        new.N = self.N + Nn
        new.csr_ptr = new.csr_ent = new.tiles = new.groups = new.sorted = new.rows = None
        new._bufs, new._tables = {}, {}
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            greedy = torch.empty((Nn, d + 1), dtype=torch.int16, device=dev)
            rank = torch.empty((Nn, d + 1), dtype=torch.int8, device=dev)
            replay = torch.empty((Nn, d + 1, 2), dtype=torch.int32, device=dev)
            flags = torch.zeros(1, dtype=torch.int32, device=dev)
            M_add = 0
            total = Nn * (d + 1)
            cap = int(lib.sgp_hash_capacity(self.M + total))
            new.hash_capacity = cap
            table = torch.full((cap,), -1, dtype=torch.int64, device=dev)
            check(lib.sgp_hash_seed(_ptr(self.keys), self.M, d, _ptr(table), cap, _ptr(flags), st))
            if Nn > 0:
                check(lib.sgp_build_points(_ptr(x_new), Nn, d, x_new.stride(0), _fp(self.scale), _ptr(greedy),
                                           _ptr(rank), _ptr(replay), _ptr(flags), st))
                slot_of = torch.empty(total, dtype=torch.int32, device=dev)
                check(lib.sgp_hash_extend(_ptr(greedy), _ptr(rank), Nn, d, _ptr(self.keys), self.M, _ptr(table), cap,
                                          _ptr(slot_of), _ptr(flags), st))
                ws_bytes = int(lib.sgp_number_workspace_bytes(Nn, d))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                m_add, fl = C.c_int64(0), C.c_int32(0)
                check(lib.sgp_count_extension(_ptr(table), cap, _ptr(slot_of), Nn, d, _ptr(ws), ws_bytes, _ptr(flags),
                                              C.byref(m_add), C.byref(fl), st))
                M_add = int(m_add.value)
            new.M = self.M + M_add
            new._covers_all_rows = getattr(self, "_covers_all_rows", False)
            new.keys = torch.empty((new.M, d), dtype=torch.int16, device=dev)
            new.keys[:self.M].copy_(self.keys)
            if Nn > 0:
                check(lib.sgp_number_extension(_ptr(table), cap, _ptr(slot_of), _ptr(greedy), _ptr(rank), Nn, d,
                                               _ptr(ws), self.M, M_add, _ptr(replay), _ptr(new.keys), st))
                del slot_of, ws
            new.replay = torch.cat([self.replay, replay], dim=0)
            if self.greedy is not None and self.rank is not None:
                new.greedy, new.rank = torch.cat([self.greedy, greedy], dim=0), torch.cat([self.rank, rank], dim=0)
            else:
                new.greedy = new.rank = None
            new.nbr = torch.empty((d + 1, 0, 2 * r), dtype=torch.int32, device=dev)
            lazy = int(lazy_tables) > 0 and (build_groups or build_rows)
            if new.N > 0 and new.M > 0:
                new._build_derived(table, build_nbr=build_nbr or lazy, build_csr=build_csr,
                                   build_groups=build_groups and not lazy, group_axes=group_axes,
                                   group_rows=group_rows, build_rows=build_rows and not lazy)
                if lazy:
                    new._lazy = {"after": int(lazy_tables), "calls": 0, "build_groups": build_groups,
                                 "group_axes": group_axes, "group_rows": group_rows, "build_rows": build_rows}
            del table
        return new

    def _count_product(self) -> None:
        """Build the postponed blur groups / row-sorted entries once enough products have been asked for."""
        lazy = getattr(self, "_lazy", None)
        if lazy is None:
            return
        lazy["calls"] += 1
        if lazy["calls"] <= lazy["after"]:
            return
        self._lazy = None
        with torch.cuda.device(self.device):
            if lazy["build_groups"] and self.order >= 1 and self.M > 0:
                self._build_groups(lazy["group_axes"], lazy["group_rows"])
            if lazy["build_rows"] and self.M > 0 and self.rows is None:
                self._build_rows()
        self._tables = {}

    @classmethod
    def from_arrays(cls, coeffs, replay: torch.Tensor, keys: torch.Tensor, nbr: Optional[torch.Tensor],
                    build_csr: bool = False, build_tiles: bool = False, tile_points: int = 256,
                    build_groups: bool = True, group_axes: Optional[int] = None, group_rows: int = 512,
                    exact: bool = False, table: Optional[torch.Tensor] = None, build_nbr: bool = True,
                    build_rows: bool = True) -> "Lattice":
        """Wrap lattice arrays that were built elsewhere: received by ``distributed.broadcast_lattice`` (``nbr`` given),
        or merged from per-rank builds (``nbr=None`` and ``table`` = the key -> index hash table of ``keys``, from which
        the neighbour table -- or, with ``build_nbr=False``, the blur groups directly -- are made)."""
        self = object.__new__(cls)
        self.device = replay.device
        self.N, self.d = int(replay.shape[0]), int(replay.shape[1]) - 1
        self.coeffs = _coeffs_np(coeffs)
        self.order = self.coeffs.shape[0] // 2
        self.var = stencil_variance(self.coeffs)
        self.scale = scale_factors(self.d, self.var)
        self.M = int(keys.shape[0])
        if nbr is None and table is None:
            raise ValueError("from_arrays needs the neighbour table or the key hash table")
        if nbr is not None and tuple(nbr.shape) != (self.d + 1, self.M, 2 * self.order):
            raise ValueError(f"nbr shape {tuple(nbr.shape)} does not match (d+1, M, 2r)")
        self.replay, self.keys = replay.contiguous(), keys.contiguous()
        self.nbr = nbr.contiguous() if nbr is not None else None
        self.greedy = self.rank = None
        self.csr_ptr = self.csr_ent = None
        self.tiles = None
        self.groups = None
        self.sorted = None      # the locality order needs greedy, which does not travel with the arrays
        self.rows = None
        self.exact = bool(exact)
        self.hash_capacity = 0 if table is None else int(table.numel())
        self._bufs = {}
        self._tables = {}
        if self.M > 0:
            with torch.cuda.device(self.device):
                if nbr is None:
                    self._build_derived(table, build_nbr=build_nbr, build_csr=build_csr and self.N > 0,
                                        build_groups=build_groups, group_axes=group_axes, group_rows=group_rows,
                                        build_rows=build_rows and self.N > 0)
                elif self.N > 0:
                    if build_csr:
                        self._build_csr()
                    if build_groups and self.order >= 1:
                        self._build_groups(group_axes, group_rows)
                    if build_rows and self.rows is None:
                        self._build_rows()
        return self

    def _launch_groups_hinted(self, group_axes, group_rows, table=None):
        """Blur groups with the axis ranges of the last build of this (d, order, rows) -- no probing, no synchronisation:
        returns the pending state for ``_finish_groups_hinted`` or None when there is no usable hint."""
        import os
        if not (group_axes is None or int(group_axes) <= 0) or os.environ.get("SGP_GROUP_FAST", "1") == "0":
            return None
        lib = _capi.lib()
        dev, d, M, r = self.device, self.d, self.M, self.order
        rows_limit = max(1, min(int(group_rows), 1024))
        hint = _GROUP_HINTS.get((d, r, rows_limit))
        if not isinstance(hint, dict) or sum(hint["lengths"]) != d + 1:
            return None
        lengths, mcs = hint["lengths"], hint["max_class"]
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        ws_bytes = int(lib.sgp_group_workspace_bytes(M))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        results = torch.zeros((len(lengths), 8), dtype=torch.int32, device=dev)
        groups, keep, prev_pos, j0 = [], [ws, results, table], None, 0
        for gi, (ln, mc) in enumerate(zip(lengths, mcs)):
            j1 = j0 + ln
            bound_class = min(rows_limit - 1, mc + max(64, mc // 4))    # room for the classes to grow since the last build
            max_batches = int(lib.sgp_group_max_batches(M, rows_limit, bound_class))
            if max_batches > max(M // 8, 1 << 16):      # (batch_begin would be out of proportion: probe the slow way)
                return None
            order_of = torch.empty(M, dtype=torch.int32, device=dev)
            pos = torch.empty(M, dtype=torch.int32, device=dev)
            cstart = torch.empty(M, dtype=torch.int32, device=dev)
            g = {"j0": j0, "j1": j1, "batch_begin": torch.empty(max_batches + 1, dtype=torch.int32, device=dev),
                 "src": torch.empty(M, dtype=torch.int32, device=dev),
                 "lnb": torch.empty((M, j1 - j0, 2 * r), dtype=torch.int16, device=dev), "max_batches": max_batches}
            check(lib.sgp_group_prepare_async(_ptr(self.keys), M, d, j0, j1, _ptr(order_of), _ptr(pos), _ptr(cstart), _ptr(ws),
                                              ws_bytes, _ptr(results[gi]), C.c_void_p(main.cuda_stream)))
            side.wait_stream(main)
            check(lib.sgp_group_finalize_async(_ptr(self.nbr if table is None else None), _ptr(self.keys), d, _ptr(table),
                                               0 if table is None else int(table.numel()), M, r, j0, j1, _ptr(order_of),
                                               _ptr(pos), _ptr(cstart), _ptr(prev_pos), rows_limit, max_batches,
                                               _ptr(g["batch_begin"]), _ptr(g["src"]), _ptr(g["lnb"]), _ptr(results[gi]),
                                               C.c_void_p(side.cuda_stream)))
            groups.append(g)
            keep += [order_of, cstart, pos]
            g["pos"] = pos
            prev_pos = pos
            j0 = j1
        return {"groups": groups, "results": results, "keep": keep, "rows_limit": rows_limit, "side": side}

    def _finish_groups_hinted(self, pending) -> bool:
        """The one synchronisation of the hinted group build: read every group's figures, accept or reject them all."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(pending["side"])
        res = pending["results"].cpu().numpy().astype("int64") & 0xFFFFFFFF        # synchronises
        rows_limit, groups = pending["rows_limit"], pending["groups"]
        for g, r_ in zip(groups, res):
            if not (0 < r_[0] <= rows_limit and r_[3] == 0 and r_[4] == 0 and 1 <= r_[1] <= g["max_batches"]):
                return False
        for g, r_ in zip(groups, res):
            g["max_class"], g["n_batches"], g["rows_cap"] = int(r_[0]), int(r_[1]), int(r_[2])
        arr = (_capi.BlurGroup * len(groups))()
        for k, g in enumerate(groups):
            arr[k] = _capi.BlurGroup(g["j0"], g["j1"], g["rows_cap"], 1024 if rows_limit > 512 else 512, g["n_batches"],
                                     g["batch_begin"].data_ptr(), g["src"].data_ptr(), g["lnb"].data_ptr())
        self.groups = {"list": groups, "array": arr, "final_pos": groups[-1]["pos"]}
        _GROUP_HINTS[(self.d, self.order, rows_limit)] = {"lengths": [g["j1"] - g["j0"] for g in groups],
                                                          "max_class": [g["max_class"] for g in groups]}
        return True

    def _build_groups(self, group_axes: Optional[int] = None, group_rows: int = 512, table=None) -> None:
        """Blur groups (csrc/sgp_groups.cu): cover axes 0..d with ranges of consecutive axes whose classes fit
        ``group_rows`` rows of one CTA.  ``group_axes=None``: every range is made as long as it can be (3 axes at the
        metric configuration, 10 on the sparse d = 18 lattice); an integer fixes the length.  A range is shortened until
        its largest class fits; if even a single axis does not fit (a long 1-D line), no groups are built and the
        per-axis blur is used."""
        lib = _capi.lib()
        dev, d, M, r = self.device, self.d, self.M, self.order
        rows_limit = max(1, min(int(group_rows), 1024))   # a CTA of 256 threads holds 512 rows, of 512 threads 1024
        st = _stream_ptr(dev)
        ws_bytes = int(lib.sgp_group_workspace_bytes(M))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        groups, keep = [], []
        prev_pos = None
        j0 = 0
        def prepare(j0, j1):
            order_of = torch.empty(M, dtype=torch.int32, device=dev)
            pos = torch.empty(M, dtype=torch.int32, device=dev)
            cstart = torch.empty(M, dtype=torch.int32, device=dev)
            mx = C.c_int64(0)
            check(lib.sgp_group_prepare(_ptr(self.keys), M, d, j0, j1, _ptr(order_of), _ptr(pos), _ptr(cstart),
                                        _ptr(ws), ws_bytes, C.byref(mx), st))
            return order_of, pos, cstart, mx

        adaptive = group_axes is None or int(group_axes) <= 0
        # the ranges the last lattice of this shape ended up with: successive hyper-parameter steps build almost the
        # same lattice, and on sparse high-dimensional lattices finding a 10-axis range one axis at a time costs more
        # than everything else in the build (16 sorts at d = 18)
        hint_key = (d, r, rows_limit)
        hint = _GROUP_HINTS.get(hint_key) if adaptive else None
        hints = hint["lengths"] if isinstance(hint, dict) else []
        while j0 <= d:
            start = hints[len(groups)] if len(groups) < len(hints) else 3
            j1 = min(j0 + (start if adaptive else max(1, int(group_axes))), d + 1)
            order_of, pos, cstart, mx = prepare(j0, j1)
            while mx.value > rows_limit and j1 - j0 > 1:      # shorten the range until its largest class fits
                j1 -= 1
                order_of, pos, cstart, mx = prepare(j0, j1)
            hinted = adaptive and len(groups) < len(hints) and j1 - j0 == hints[len(groups)]
            if adaptive and not hinted:   # (a range that repeats the last build's length is not probed for one more axis)
                # ... or lengthen it while it still does and the batch-local neighbour table stays at most 48 bytes a
                # row (it shares the CTA's shared memory with the values: longer tables cost occupancy)
                max_axes = max(3, 48 // (4 * r))
                # (one more axis at least doubles the largest class of a lattice this dense: do not pay for a trial that
                # cannot fit)
                while mx.value * 2 <= rows_limit and j1 <= d and j1 - j0 < max_axes:
                    trial = prepare(j0, j1 + 1)
                    if trial[3].value > rows_limit:
                        break
                    order_of, pos, cstart, mx = trial
                    j1 += 1
            if mx.value > rows_limit:
                self.groups = None
                return
            max_batches = int(lib.sgp_group_max_batches(M, rows_limit, int(mx.value)))
            g = {
                "j0": j0, "j1": j1, "max_class": int(mx.value),
                "batch_begin": torch.empty(max_batches + 1, dtype=torch.int32, device=dev),
                "src": torch.empty(M, dtype=torch.int32, device=dev),
                "lnb": torch.empty((M, j1 - j0, 2 * r), dtype=torch.int16, device=dev),
            }
            rows, nb = C.c_int32(0), C.c_int64(0)
            check(lib.sgp_group_finalize(_ptr(self.nbr if table is None else None), _ptr(self.keys), d, _ptr(table),
                                         0 if table is None else int(table.numel()), M, r, j0, j1, _ptr(order_of),
                                         _ptr(pos), _ptr(cstart),
                                         _ptr(prev_pos), rows_limit, max_batches, _ptr(g["batch_begin"]), _ptr(g["src"]),
                                         _ptr(g["lnb"]), _ptr(ws), ws_bytes, C.byref(nb), C.byref(rows), st))
            g["n_batches"] = int(nb.value)
            g["rows_cap"] = int(rows.value)
            groups.append(g)
            prev_pos = pos
            keep.append(pos)
            j0 = j1
        arr = (_capi.BlurGroup * len(groups))()
        for k, g in enumerate(groups):
            arr[k] = _capi.BlurGroup(g["j0"], g["j1"], g["rows_cap"], 1024 if rows_limit > 512 else 512, g["n_batches"],
                                     g["batch_begin"].data_ptr(),
                                     g["src"].data_ptr(), g["lnb"].data_ptr())
        self.groups = {"list": groups, "array": arr, "final_pos": prev_pos}
        if adaptive:
            _GROUP_HINTS[hint_key] = {"lengths": [g["j1"] - g["j0"] for g in groups],
                                      "max_class": [g["max_class"] for g in groups]}

    def _build_rows(self) -> None:
        """Point-vertices sorted by lattice row for the segmented-gather splat (csrc/sgp_tiles.cu, sgp_build_rowsorted):
        ``ent {point | row-start flag, weight}`` and ``seg_row`` (lattice row of every fourth entry).  The encoding needs
        an entry in every lattice row: a lattice wrapped around a subset of the points (``from_arrays`` under point
        sharding) gets one weightless filler entry per row."""
        lib = _capi.lib()
        dev, N, d, M = self.device, self.N, self.d, self.M
        st = _stream_ptr(dev)
        total = N * (d + 1)

        def build(fill):
            n = int(lib.sgp_rowsort_padded(N, d, fill))
            ws_bytes = int(lib.sgp_rowsort_workspace_bytes(N, d, fill))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            ent = torch.empty((n, 2), dtype=torch.int32, device=dev)
            seg_row = torch.empty(n // 4, dtype=torch.int32, device=dev)
            check(lib.sgp_build_rowsorted(_ptr(self.replay), N, d, M, fill, _ptr(ent), None, _ptr(seg_row), _ptr(ws),
                                          ws_bytes, st))
            return {"ent": ent, "seg_row": seg_row, "n": n, "entries": total + fill}

        rows = build(0)
        if not getattr(self, "_covers_all_rows", False):
            # wrapped arrays: do the points touch every row?  (row starts = flags + the first entry)
            starts = int((rows["ent"][:, 0] < 0).sum()) + 1      # (padding entries carry no flag; the storage is interleaved)
            if starts != M:
                rows = build(M)
        self.rows = rows

    def _sort_points(self) -> None:
        """Locality order of the points (csrc/sgp_tiles.cu, sgp_sort_points); the replay tables re-ordered with it are
        made on first use by ``_table``."""
        lib = _capi.lib()
        dev, N, d = self.device, self.N, self.d
        st = _stream_ptr(dev)
        ws_bytes = int(lib.sgp_sort_points_workspace_bytes(N))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        perm = torch.empty(N, dtype=torch.int32, device=dev)
        check(lib.sgp_sort_points(_ptr(self.greedy), N, d, _ptr(perm), _ptr(ws), ws_bytes, st))
        self.sorted = {"perm": perm}

    def _build_tiles(self, tile_points: int = 256) -> None:
        """Locality tiles for the shared-memory staged splat / slice (csrc/sgp_tiles.cu): the points in locality
        order are cut into tiles of ``tile_points``; per tile, the distinct lattice rows it touches (its dictionary)
        and its point-vertices grouped by row into segments, the segments cut into pieces of at most 8 entries."""
        lib = _capi.lib()
        dev, N, d, M = self.device, self.N, self.d, self.M
        if self.sorted is None:
            self._sort_points()
        perm = self.sorted["perm"]
        total = N * (d + 1)
        T = max(8, int(tile_points) // 8 * 8)
        while T > 8 and T * (d + 1) > 65535:
            T //= 2
        n_tiles = (N + T - 1) // T
        st = _stream_ptr(dev)
        ws_bytes = int(lib.sgp_tiles_workspace_bytes(N, d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        S = C.c_int64(0)
        check(lib.sgp_tiles_prepare(_ptr(self.replay), N, d, M, T, _ptr(perm), _ptr(ws), ws_bytes, C.byref(S), st))
        S = int(S.value)
        seg_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
        t = {
            "T": T, "S": S,
            "seg_row": torch.empty(S, dtype=torch.int32, device=dev),
            "seg_ent": torch.empty((total, 2), dtype=torch.int32, device=dev),
            "tile_seg_ptr": torch.empty(n_tiles + 1, dtype=torch.int32, device=dev),
            "lidx": torch.zeros(total + 8, dtype=torch.int16, device=dev),   # padded: copied in 4-byte pieces
            "tile_w": torch.empty(total, dtype=torch.float32, device=dev),
        }
        mx = C.c_int32(0)
        check(lib.sgp_tiles_finalize(_ptr(self.replay), _ptr(perm), N, d, T, S, _ptr(ws), ws_bytes, _ptr(seg_ptr),
                                     _ptr(t["seg_row"]), _ptr(t["seg_ent"]), _ptr(t["tile_seg_ptr"]), _ptr(t["lidx"]),
                                     _ptr(t["tile_w"]), C.byref(mx), st))
        t["max_dict"] = int(mx.value)
        del ws
        # splat pieces: every segment cut into runs of at most 8 entries (index plumbing of the build, torch ops)
        PIECE = 8
        sp = seg_ptr.long()
        lens = sp[1:] - sp[:-1]
        npc = (lens + (PIECE - 1)) // PIECE
        first = torch.cumsum(npc, 0) - npc
        P = int(npc.sum().item())
        seg_of_piece = torch.repeat_interleave(torch.arange(S, device=dev), npc, output_size=P)
        k_in_seg = torch.arange(P, device=dev) - first[seg_of_piece]
        start = sp[:-1][seg_of_piece] + PIECE * k_in_seg
        piece_ptr = torch.empty(P + 1, dtype=torch.int32, device=dev)
        piece_ptr[:P] = start.to(torch.int32)
        piece_ptr[P] = total
        t["P"] = P
        t["piece_ptr"] = piece_ptr
        t["piece_row"] = t["seg_row"][seg_of_piece].contiguous()
        tsp = t["tile_seg_ptr"].long()
        first_ext = torch.cat([first, first.new_tensor([P])])
        t["tile_piece_ptr"] = first_ext[tsp].to(torch.int32).contiguous()
        # the slice after a blur-group chain reads the lattice values in the last stage's order
        t["seg_row_out"] = t["seg_row"]
        if self.groups is not None:
            t["seg_row_out"] = self.groups["final_pos"][t["seg_row"].long()].contiguous()
        self.tiles = t

    def _tiles_view(self, final: bool = False) -> TilesView:
        t = self.tiles
        rows = t["seg_row_out"] if final else t["seg_row"]
        return TilesView(self.N, self.M, t["S"], t["P"], self.d, t["T"], t["max_dict"], 0,
                         self.sorted["perm"].data_ptr(), t["tile_seg_ptr"].data_ptr(), rows.data_ptr(),
                         t["lidx"].data_ptr(), t["tile_w"].data_ptr(), t["seg_ent"].data_ptr(),
                         t["tile_piece_ptr"].data_ptr(), t["piece_ptr"].data_ptr(), t["piece_row"].data_ptr())

    def _build_csr(self) -> None:
        """Row starts for the ordered-gather splat (``mode=2``): the row-sorted entries are already in the reference's
        accumulation order (stable sort), so the CSR form is those entries plus a pointer array."""
        if self.rows is None:
            self._build_rows()
        total = self.rows["entries"]
        # row-start flags; the entries are stored interleaved in groups of 64 (csrc/sgp_common.cuh, sgp_entry_index):
        # storage position j holds row-sorted entry  group*64 + segment*8 + piece*2 + half
        j = torch.nonzero(self.rows["ent"][:, 0] < 0).flatten()
        within = j % 64
        lin = (j // 64) * 64 + ((within % 16) // 2) * 8 + (within // 16) * 2 + (within % 2)
        starts = torch.sort(lin).values.to(torch.int32)
        ends = torch.tensor([total], dtype=torch.int32, device=self.device)
        self.csr_ptr = torch.cat([torch.zeros(1, dtype=torch.int32, device=self.device), starts, ends]).contiguous()
        if self.csr_ptr.numel() != self.M + 1:
            raise RuntimeError("row-sorted entries do not cover every lattice row")
        self.csr_ent = self.rows["ent"]

    # ---- structure accessors (reference numbering) -------------------------------------------
    @property
    def offsets(self) -> torch.Tensor:
        """int32 ``[N, d+1]``: lattice index of every simplex vertex (reference ``replay[].offset / vd``)."""
        return self.replay[..., 0]

    @property
    def weights(self) -> torch.Tensor:
        """fp32 ``[N, d+1]``: barycentric weight of every simplex vertex (reference ``replay[].weight``)."""
        return self.replay[..., 1].view(torch.float32)

    def _table(self, sorted: bool, final: bool, transposed: bool = False) -> torch.Tensor:
        """Internal replay table ``[N, d+1, 2]``, built on first use: rows in the locality order when ``sorted``;
        lattice indices mapped to the order the last blur-group stage leaves the lattice values in when ``final``;
        ``transposed``: ``[d+1, N, 2]`` -- the layout for narrow rows (one or two channel chunks), where the lanes of a
        warp are different points and read one contiguous run per vertex instead of 72-byte strides."""
        key = (bool(sorted), bool(final), bool(transposed))
        if key == (False, False, False):
            return self.replay
        t = self._tables.get(key)
        if t is None and transposed:
            perm = self.sorted["perm"] if sorted else None
            pos = self.groups["final_pos"] if final else None
            t = torch.empty((self.d + 1, self.N, 2), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                check(_capi.lib().sgp_permute_replay(_ptr(self.replay), _ptr(perm), _ptr(pos), self.N, self.d, 1, _ptr(t),
                                                     _stream_ptr(self.device)))
            self._tables[key] = t
        if t is None:
            perm = self.sorted["perm"] if sorted else None
            pos = self.groups["final_pos"] if final else None
            # an odd number of vertices per point gets one filler slot: every point of the table then starts 16-byte
            # aligned and the ring slice reads its entries in pairs (SGP_REPLAY_PAD=0 keeps the dense table)
            import os
            stride = self.d + 1
            if final and not sorted and stride % 2 == 1 and os.environ.get("SGP_REPLAY_PAD", "1") != "0":
                stride += 1
            t = torch.empty((self.N, stride, 2), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                if stride == self.d + 1:
                    check(_capi.lib().sgp_permute_replay(_ptr(self.replay), _ptr(perm), _ptr(pos), self.N, self.d, 0, _ptr(t),
                                                         _stream_ptr(self.device)))
                else:
                    check(_capi.lib().sgp_permute_replay_padded(_ptr(self.replay), _ptr(perm), _ptr(pos), self.N, self.d,
                                                                stride, _ptr(t), _stream_ptr(self.device)))
            self._tables[key] = t
        return t

    def _view(self, replay: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None,
              exact: Optional[bool] = None, transposed: bool = False) -> LatticeView:
        exact = self.exact if exact is None else exact
        rp = self.replay if replay is None else replay
        stride = int(rp.shape[1]) if (not transposed and rp.dim() == 3 and int(rp.shape[1]) != self.d + 1) else 0
        return LatticeView(self.N, self.M, self.d, self.order, rp.data_ptr(),
                           self.nbr.data_ptr() if (self.nbr is not None and self.nbr.numel()) else 0,
                           self.csr_ptr.data_ptr() if self.csr_ptr is not None else 0,
                           self.csr_ent.data_ptr() if self.csr_ent is not None else 0,
                           0 if perm is None else perm.data_ptr(), 0 if exact else 1, 1 if transposed else 0, stride, 0)

    def _slice_view(self, Lv: int, final: bool, exact: Optional[bool] = None) -> LatticeView:
        """The view the slice of an ``Lv``-channel product reads: the dense / padded table for wide rows (several lanes
        per point: TMA-ring slice), the transposed table for narrow ones (SGP_SLICE_TRANSPOSED_MAX_L, default 8: measured 24 / 30 / 36 / 39 us against
        29 / 33 / 40 / 46 us at 1 / 2 / 4 / 8 channels, metric shape)."""
        import os
        narrow = int(Lv) <= int(os.environ.get("SGP_SLICE_TRANSPOSED_MAX_L", "8"))
        if narrow:
            return self._view(self._table(False, final, True), None, exact, transposed=True)
        return self._view(self._table(False, final), None, exact)

    def _scratch(self, L: int):
        key = int(L)
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) >= 2:
                self._bufs.clear()
            b = (torch.empty((max(self.M, 1), L), dtype=torch.float32, device=self.device),
                 torch.empty((max(self.M, 1), L), dtype=torch.float32, device=self.device))
            self._bufs[key] = b
        return b

    def lattice_width(self, L: int) -> int:
        """Channels of the lattice value rows for an ``L``-column block on the row-sorted chain: ``L`` rounded up to a
        multiple of four (16-byte vectors between splat and slice), and 16 instead of 12 on large dense lattices --
        48-byte rows straddle 128-byte lines, so every fourth row gather costs two L1 wavefronts; on 64-byte rows the
        spare lanes idle and splat, blur and slice are all faster (config A, 12 columns: 181 -> 170 us per MVM).
        ``SGP_LV16=0`` turns that off, ``SGP_LV16=2`` applies it whatever the lattice (tests)."""
        if L <= 4:
            return L
        Lv = (L + 3) // 4 * 4
        if Lv == 12 and self.rows is not None:
            import os
            e = os.environ.get("SGP_LV16", "1")
            if e == "2" or (e != "0" and self.rows["n"] >= 8 * self.M and self.rows["n"] >= (1 << 21)):
                Lv = 16
        return Lv

    def _pads_ragged_src(self) -> bool:
        """Copy a ragged right-hand-side block into a zero-padded one in front of the splat?  Pays once the splat is
        bound by its gathers (SGP_PAD_SRC=0/1 forces)."""
        import os
        e = os.environ.get("SGP_PAD_SRC")
        if e is not None:
            return e != "0"
        return self.N * (self.d + 1) >= (1 << 20)

    def _src_pad(self, Lv: int) -> torch.Tensor:
        p = self._bufs.get(("pad", Lv))
        if p is None:
            for k in [k for k in self._bufs if isinstance(k, tuple)]:
                del self._bufs[k]
            p = torch.zeros((self.N, Lv), dtype=torch.float32, device=self.device)
            self._bufs[("pad", Lv)] = p
        return p

    def _check_src(self, src: torch.Tensor) -> torch.Tensor:
        if src.dim() != 2 or src.shape[0] != self.N:
            raise ValueError(f"Incompatible shapes {tuple(src.shape)}, and {(self.N, self.d)}")
        if src.device != self.device or src.dtype != torch.float32:
            raise TypeError("src must be float32 on the lattice's device")
        if src.stride(1) != 1 and src.shape[1] > 1:
            src = src.contiguous()
        return src

    # ---- stages, exposed separately for parity tests ---------------------------------------------
    # Stage methods take and return lattice values in LATTICE-INDEX order (the reference's numbering) and default to
    # the reference's arithmetic (exact=True); `mvm` is the production path and defaults to the lattice's setting.
    def splat(self, src: torch.Tensor, mode: int = _capi.SGP_SPLAT_AUTO, sorted: bool = False) -> torch.Tensor:
        src = self._check_src(src)
        L = int(src.shape[1])
        values = torch.empty((self.M, L), dtype=torch.float32, device=self.device)
        if self.M == 0 or L == 0:
            return values
        with torch.cuda.device(self.device):
            if mode == _capi.MODE_TILES:
                tv = self._tiles_view()
                check(_capi.lib().sgp_splat_tiles(C.byref(tv), _ptr(src), src.stride(0), L, _ptr(values),
                                                  _stream_ptr(self.device)))
            elif mode == _capi.MODE_ROWS:
                check(_capi.lib().sgp_splat_rows(_ptr(self.rows["ent"]), _ptr(self.rows["seg_row"]), self.rows["n"],
                                                 self.N, self.M, _ptr(src), src.stride(0), L, _ptr(values), L,
                                                 _stream_ptr(self.device)))
            else:
                if mode == _capi.MODE_AUTO:
                    mode = _capi.MODE_GATHER if self.csr_ptr is not None else _capi.MODE_ATOMIC
                v = self._view()
                if sorted and mode == _capi.MODE_ATOMIC:
                    v = self._view(self._table(True, False), self.sorted["perm"])
                check(_capi.lib().sgp_splat(C.byref(v), _ptr(src), src.stride(0), L, _ptr(values), mode,
                                            _stream_ptr(self.device)))
        return values

    def blur(self, values: torch.Tensor, coeffs=None, groups: bool = False, exact: bool = True) -> torch.Tensor:
        """Blurred lattice values in lattice-index order.  ``groups=True`` runs the shared-memory group chain (whose
        output order is its last stage's) and permutes the result back, for comparison with the per-axis path."""
        c = self.coeffs if coeffs is None else _coeffs_np(coeffs)
        L = int(values.shape[1])
        if self.M == 0 or L == 0:
            return values.clone()
        buf0 = values.contiguous().clone()
        buf1 = torch.empty_like(buf0)
        where = C.c_int(0)
        with torch.cuda.device(self.device):
            if groups:
                if self.groups is None:
                    raise RuntimeError("blur groups were not built for this lattice")
                arr = self.groups["array"]
                check(_capi.lib().sgp_blur_groups(arr, len(arr), self.M, self.order, _fp(c), c.shape[0], L, _ptr(buf0),
                                                  _ptr(buf1), C.byref(where), 0 if exact else 1,
                                                  _stream_ptr(self.device)))
                res = buf1 if where.value else buf0
                return res[self.groups["final_pos"].long()]
            v = self._view(exact=exact)
            check(_capi.lib().sgp_blur(C.byref(v), _fp(c), c.shape[0], L, _ptr(buf0), _ptr(buf1), C.byref(where),
                                       _stream_ptr(self.device)))
        return buf1 if where.value else buf0

    def slice(self, values: torch.Tensor, mode: int = _capi.MODE_AUTO, sorted: bool = False,
              exact: bool = True) -> torch.Tensor:
        L = int(values.shape[1])
        out = torch.empty((self.N, L), dtype=torch.float32, device=self.device)
        if self.N == 0 or L == 0:
            return out
        values = values.contiguous()
        with torch.cuda.device(self.device):
            if mode == _capi.MODE_TILES:
                tv = self._tiles_view()
                check(_capi.lib().sgp_slice_tiles(C.byref(tv), _ptr(values), L, _ptr(out), out.stride(0),
                                                  0 if exact else 1, _stream_ptr(self.device)))
            else:
                v = self._view(self._table(True, False), self.sorted["perm"], exact) if sorted \
                    else self._view(exact=exact)
                check(_capi.lib().sgp_slice(C.byref(v), _ptr(values), L, _ptr(out), out.stride(0), L,
                                            _stream_ptr(self.device)))
        return out

    # ---- the MVM ------------------------------------------------------------------------------------
    def mvm(self, src: torch.Tensor, out: Optional[torch.Tensor] = None, coeffs=None,
            mode: int = _capi.SGP_SPLAT_AUTO, blur: str = "auto", sorted: Optional[bool] = None,
            exact: Optional[bool] = None, after_splat=None, scratch=None, zero_flags: int = 0, cg=None) -> torch.Tensor:
        """``out[N, L] = slice(blur(splat(src[N, L])))`` on the built lattice.

        ``mode``   splat form: 0 auto (row-sorted segmented gather when built, else atomic scatter), 1 atomic scatter
                   (``red.global.add.v4.f32`` per point-vertex), 2 gather in the reference's accumulation order
                   (deterministic; needs ``build_csr=True``), 3 locality tiles (needs ``build_tiles=True``; implies the
                   tile slice), 4 row-sorted segmented gather (``build_rows=True``).
        ``blur``   "groups" (several axes per launch through shared memory), "axis" (one launch per axis) or "auto"
                   (groups when they were built).
        ``sorted`` walk the points in the locality order in splat and slice (needs ``sort_points=True``; default off:
                   measured slower than the input order on B200).
        ``exact``  the reference's arithmetic (one rounding per product and sum, one division per slice term) instead
                   of fused multiply-adds; with ``mode=2`` the result is then bit-identical to the reference's.
                   Default: the lattice's ``exact`` attribute (False).
        ``after_splat`` optional callable invoked with the splatted lattice values ``[M, L]`` (in place) before the
                   blur: the exchange step of point sharding (an all-reduce over the ranks' partial splats).
        ``scratch`` a private pair of ``[M, ceil4(L)]`` work buffers instead of the lattice's shared ones (a captured CUDA
                   graph owns its pair: its nodes must not point into buffers the lattice may free or reuse).
        ``zero_flags`` (production chain only, private scratch): 1 = ``scratch[0]`` holds zeros on entry, 2 = leave it
                   zeroed on exit, overlapped with the slice (``sgp_mvm_rows_groups_ex``; what ``capture`` uses).
        ``cg``     ``(s, noise, pAp, scratch)`` device tensors: also run the sweep that follows the product in a CG
                   iteration -- ``out = s * K src + noise * src``, ``pAp[l] = sum_n src * out`` (``sgp_cg_apply``) -- folded
                   into the slice's epilogue on the production chain (``sgp_mvm_rows_groups_cg``), as its own launch
                   otherwise.  Needs ``L % 4 == 0`` or ``L <= 4`` (no channel padding) and contiguous blocks."""
        src = self._check_src(src)
        L = int(src.shape[1])
        c = self.coeffs if coeffs is None else _coeffs_np(coeffs)
        if c.shape[0] != 2 * self.order + 1:
            raise ValueError("stencil length does not match the order this lattice was built for")
        if out is None:
            out = torch.empty((self.N, L), dtype=torch.float32, device=self.device)
        elif (tuple(out.shape) != (self.N, L) or out.dtype != torch.float32 or out.device != self.device
              or (L > 1 and out.stride(1) != 1)):
            raise ValueError(f"out must be a float32 [{self.N}, {L}] tensor on {self.device} with unit column stride")
        if self.N == 0 or L == 0:
            return out
        self._count_product()
        exact = self.exact if exact is None else bool(exact)
        lib, st = _capi.lib(), _stream_ptr(self.device)
        use_tiles = mode == _capi.MODE_TILES
        if use_tiles and self.tiles is None:
            raise RuntimeError("tiles were not built for this lattice (build_tiles=True)")
        if mode == _capi.MODE_AUTO:
            mode = _capi.MODE_ROWS if self.rows is not None else _capi.MODE_ATOMIC
        if mode == _capi.MODE_ROWS and self.rows is None:
            raise RuntimeError("the row-sorted entries were not built for this lattice (build_rows=True)")
        # Lattice rows are padded to a multiple of 4 channels so that everything between splat and slice moves
        # 16-byte vectors (L = 11, the CG block of a training step: 301 us on the scalar path at the metric shape);
        # the row-sorted splat and the slice read / write the caller's ragged rows channel by channel.
        Lv = self.lattice_width(L) if mode == _capi.MODE_ROWS else L
        buf0, buf1 = self._scratch(Lv) if scratch is None else scratch[:2]
        if tuple(buf0.shape) != (max(self.M, 1), Lv) or tuple(buf1.shape) != (max(self.M, 1), Lv):
            raise ValueError(f"scratch buffers must be [{max(self.M, 1)}, {Lv}]")
        use_sorted = False if sorted is None else bool(sorted)   # measured slower than the input order on B200
        if use_sorted and self.sorted is None:
            raise RuntimeError("the locality order was not built for this lattice (sort_points=True)")
        use_groups = (self.groups is not None) if blur == "auto" else (blur == "groups")
        if use_groups and self.groups is None:
            raise RuntimeError("blur groups were not built for this lattice")
        if not use_groups and self.order > 0 and self.nbr is None:
            raise RuntimeError("the per-axis blur needs the neighbour table (build_nbr=True)")
        where = C.c_int(0)
        if mode == _capi.MODE_ROWS and use_groups and not use_tiles and not use_sorted and after_splat is None:
            # the production chain, one call across the C ABI
            arr = self.groups["array"]
            v_out = self._slice_view(Lv, True, exact)
            flags = int(zero_flags)
            if zero_flags and scratch is None:
                raise ValueError("zero_flags needs private scratch buffers")
            if cg is not None and (Lv == L or L % 4 == 0) and src.stride(0) == L and out.stride(0) == L:
                cs, cn, cp, cscr = cg
                with torch.cuda.device(self.device):
                    check(lib.sgp_mvm_rows_groups_cg(C.byref(v_out), _ptr(self.rows["ent"]), _ptr(self.rows["seg_row"]),
                                                     self.rows["n"], arr, len(arr), _ptr(src), src.stride(0), L, _fp(c),
                                                     c.shape[0], _ptr(out), out.stride(0), _ptr(buf0), _ptr(buf1), Lv,
                                                     flags, _ptr(cs), _ptr(cn), _ptr(cp), _ptr(cscr), st))
                return out
            with torch.cuda.device(self.device):
                if L % 4 and L > 4 and self._pads_ragged_src() and cg is None:
                    # ragged rows (L = 11: 44 bytes) can only be gathered channel by channel -- four times the L1
                    # wavefronts of 16-byte vectors, nine times per point; one coalesced copy into a zero-padded block
                    # is cheaper (config A, 11 columns: 236 -> 200 us per MVM; 181 us when the caller's block has 12)
                    Lp = (L + 3) // 4 * 4
                    pad = scratch[2] if (scratch is not None and len(scratch) > 2) else self._src_pad(Lp)
                    check(lib.sgp_pad_columns(_ptr(src), src.stride(0), L, _ptr(pad), pad.stride(0), Lp, self.N, st))
                    src, flags = pad, flags | 4      # SGP_MVM_SRC_PADDED
                if flags:
                    check(lib.sgp_mvm_rows_groups_ex(C.byref(v_out), _ptr(self.rows["ent"]), _ptr(self.rows["seg_row"]),
                                                     self.rows["n"], arr, len(arr), _ptr(src), src.stride(0), L, _fp(c),
                                                     c.shape[0], _ptr(out), out.stride(0), _ptr(buf0), _ptr(buf1), Lv,
                                                     flags, st))
                else:
                    check(lib.sgp_mvm_rows_groups(C.byref(v_out), _ptr(self.rows["ent"]), _ptr(self.rows["seg_row"]),
                                                  self.rows["n"], arr,
                                                  len(arr), _ptr(src), src.stride(0), L, _fp(c), c.shape[0], _ptr(out),
                                                  out.stride(0), _ptr(buf0), _ptr(buf1), Lv, st))
                if cg is not None:      # padded lattice rows (ragged L): the sweep as its own launch
                    if out.stride(0) != L or src.stride(0) != L:
                        raise ValueError("cg needs contiguous [N, L] blocks")
                    cs, cn, cp, cscr = cg
                    check(lib.sgp_cg_apply(_ptr(out), _ptr(src), _ptr(cs), _ptr(cn), self.N, L, _ptr(cp), _ptr(cscr), st))
            return out
        if zero_flags:
            raise ValueError("zero_flags applies to the production chain only (row-sorted splat + blur groups)")
        with torch.cuda.device(self.device):
            perm = self.sorted["perm"] if use_sorted else None
            if use_tiles:
                tv = self._tiles_view(False)
                check(lib.sgp_splat_tiles(C.byref(tv), _ptr(src), src.stride(0), L, _ptr(buf0), st))
            elif mode == _capi.MODE_ROWS:
                check(lib.sgp_splat_rows(_ptr(self.rows["ent"]), _ptr(self.rows["seg_row"]), self.rows["n"], self.N, self.M,
                                         _ptr(src), src.stride(0), L, _ptr(buf0), Lv, st))
            else:
                if mode == _capi.MODE_GATHER:
                    v_in = self._view(exact=exact)
                else:
                    v_in = self._view(self._table(use_sorted, False), perm, exact)
                check(lib.sgp_splat(C.byref(v_in), _ptr(src), src.stride(0), L, _ptr(buf0), mode, st))
            if after_splat is not None:
                after_splat(buf0[: self.M])
            if use_groups:
                arr = self.groups["array"]
                check(lib.sgp_blur_groups(arr, len(arr), self.M, self.order, _fp(c), c.shape[0], Lv, _ptr(buf0),
                                          _ptr(buf1), C.byref(where), 0 if exact else 1, st))
            else:
                vb = self._view(exact=exact)
                check(lib.sgp_blur(C.byref(vb), _fp(c), c.shape[0], Lv, _ptr(buf0), _ptr(buf1), C.byref(where), st))
            res = buf1 if where.value else buf0
            if use_tiles:
                tv = self._tiles_view(use_groups)
                check(lib.sgp_slice_tiles(C.byref(tv), _ptr(res), L, _ptr(out), out.stride(0), 0 if exact else 1, st))
            else:
                v_out = self._view(self._table(use_sorted, use_groups), perm, exact) if use_sorted \
                    else self._slice_view(Lv, use_groups, exact)
                check(lib.sgp_slice(C.byref(v_out), _ptr(res), Lv, _ptr(out), out.stride(0), L, st))
            if cg is not None:      # not the production chain: the sweep as its own launch
                if out.stride(0) != L or src.stride(0) != L:
                    raise ValueError("cg needs contiguous [N, L] blocks")
                cs, cn, cp, cscr = cg
                check(lib.sgp_cg_apply(_ptr(out), _ptr(src), _ptr(cs), _ptr(cn), self.N, L, _ptr(cp), _ptr(cscr), st))
        return out

    def capture(self, src: torch.Tensor, out: torch.Tensor, **mvm_kwargs) -> "torch.cuda.CUDAGraph":
        """CUDA graph of ``self.mvm(src, out=out, **mvm_kwargs)`` on these exact buffers (memset + splat + blur launches +
        slice: six nodes at the metric configuration).  Replaying it removes the launch gaps between the short kernels
        (215 -> 205 us per MVM on B200); the caller refills ``src`` in place and reads ``out`` after ``graph.replay()``,
        as in a CG loop with static work vectors."""
        lazy = getattr(self, "_lazy", None)
        if lazy is not None:          # a graph is for many replays: build the postponed tables now, outside the capture
            lazy["calls"] = lazy["after"]
            self._count_product()
        src = self._check_src(src)
        if out.shape != src.shape or out.dtype != torch.float32 or out.device != self.device or out.stride(1) != 1:
            raise ValueError("out must be a float32 [N, L] tensor on the lattice's device with unit column stride")
        L = int(src.shape[1])
        Lv = self.lattice_width(L)
        if mvm_kwargs.get("mode", _capi.MODE_AUTO) not in (_capi.MODE_AUTO, _capi.MODE_ROWS) or self.rows is None:
            Lv = L
        scratch = (torch.zeros((max(self.M, 1), Lv), dtype=torch.float32, device=self.device),
                   torch.empty((max(self.M, 1), Lv), dtype=torch.float32, device=self.device))
        if L % 4 and L > 4:
            scratch = scratch + (torch.zeros((self.N, (L + 3) // 4 * 4), dtype=torch.float32, device=self.device),)   # zero-padded copy of src
        # production chain on private buffers: the splat buffer is zeroed at the END of every product, next to the slice
        # (a parallel branch of the graph), instead of in front of the splat: 4-6 us off the critical path at the
        # metric shape.  SGP_GRAPH_ZERO_AFTER=0 keeps the memset in front.
        import os
        production = (mvm_kwargs.get("mode", _capi.MODE_AUTO) in (_capi.MODE_AUTO, _capi.MODE_ROWS) and self.rows is not None
                      and self.groups is not None and mvm_kwargs.get("blur", "auto") in ("auto", "groups")
                      and not mvm_kwargs.get("sorted") and mvm_kwargs.get("after_splat") is None
                      and os.environ.get("SGP_GRAPH_ZERO_AFTER", "1") != "0")
        zf = 3 if production else 0
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):     # warm-up outside the capture: lazy tables, function attributes
            self.mvm(src, out=out, scratch=scratch, zero_flags=zf, **mvm_kwargs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.mvm(src, out=out, scratch=scratch, zero_flags=zf, **mvm_kwargs)
        graph._sgp_keepalive = (scratch, src, out, self)   # the graph's nodes point into these: keep them alive with it
        return graph

    def algorithmic_bytes(self, L: int) -> int:
        """Bytes one MVM must move (SURVEY.md section 8d / BASELINE.md section 3)."""
        N, M, d, r = self.N, self.M, self.d, self.order
        return 4 * (2 * N * L + 4 * N * (d + 1) + 2 * M * L + (d + 1) * (2 * M * L + 2 * r * M))


_M_HINT = {}   # (N, d, order) -> lattice points of the last filter() call of that shape: sizes the next workspace


def _filter_call(fn, ws_fn, dev, N, d, L, order, args_before, args_after):
    """Run sgp_filter / sgp_filter_host with a workspace from PyTorch's allocator.  The workspace is sized for the
    number of lattice points the last call of this shape produced (plus a quarter), else for the worst case N(d+1) when
    that is affordable; if the lattice turns out larger the call reports M and is repeated once with the exact size."""
    total = N * (d + 1)
    hint = _M_HINT.get((N, d, order))
    if hint is not None:
        m_max = min(total, hint + hint // 4 + 1024)
    else:
        m_max = total
        free, _ = torch.cuda.mem_get_info(dev)
        while m_max > 1024 and int(ws_fn(N, d, L, order, m_max)) > free // 2:
            m_max //= 2
    m_out = C.c_int64(0)
    for attempt in range(2):
        nbytes = int(ws_fn(N, d, L, order, m_max))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        code = fn(*args_before, _ptr(ws), nbytes, m_max, C.byref(m_out), *args_after)
        if code == _capi.SGP_ENOMEM and attempt == 0 and m_out.value > m_max:
            m_max = int(m_out.value)
            del ws
            continue
        check(code)
        break
    _M_HINT[(N, d, order)] = int(m_out.value)
    return ws


def lattice_filter(src: torch.Tensor, ref: torch.Tensor, coeffs, *, device=None) -> torch.Tensor:
    """Drop-in for the reference operator ``filter(src[N,L], ref[N,d], coeffs[2r+1]) -> out[N,L]``
    (cpp/lattice.cpp:6-16 for CPU tensors, cuda/permutohedral_cuda.cpp:12-22 for CUDA tensors).

    Builds the lattice of ``ref`` and applies one MVM, as the reference does on every call
    (permutohedral.h:259-340): ONE call across the C ABI -- ``sgp_filter`` for CUDA tensors, ``sgp_filter_host`` for CPU
    tensors (the uploads and the download ride inside; pinned tensors are copied asynchronously).  The result is a fresh
    tensor with ``src``'s dtype and device; the arithmetic is fp32 whatever the input dtype (the reference's CPU filter
    is fp32-only; its CUDA twin also dispatches double).  There is no CPU path: the computation
    always runs in the CUDA kernels.
    """
    assert src.shape[0] == ref.shape[0], "Incompatible shapes {}, and {}".format(src.shape, ref.shape)
    if src.dim() != 2 or ref.dim() != 2:
        raise ValueError(f"filter: src and ref must be 2-D, got {tuple(src.shape)} and {tuple(ref.shape)}")
    if not (src.is_floating_point() and ref.is_floating_point()):
        raise TypeError("filter: floating-point tensors required")
    lib = _capi.lib()
    c = _coeffs_np(coeffs)
    N, L, d, order = int(src.shape[0]), int(src.shape[1]), int(ref.shape[1]), c.shape[0] // 2
    if not (1 <= d <= _capi.SGP_MAX_DIM):
        raise ValueError(f"d={d} outside [1, {_capi.SGP_MAX_DIM}]")
    out_dtype = src.dtype
    if src.is_cuda:
        dev = src.device
        src_d = src.detach().to(torch.float32)
        if L > 1 and src_d.stride(1) != 1:
            src_d = src_d.contiguous()
        ref_d = ref.detach().to(device=dev, dtype=torch.float32)
        if d > 1 and ref_d.stride(1) != 1:
            ref_d = ref_d.contiguous()
        out = torch.empty((N, L), dtype=torch.float32, device=dev)
        if N == 0 or L == 0:
            return out.to(out_dtype)
        with torch.cuda.device(dev):
            _filter_call(lib.sgp_filter, lib.sgp_filter_workspace_bytes, dev, N, d, L, order,
                         (_ptr(src_d), max(src_d.stride(0), L), _ptr(ref_d), max(ref_d.stride(0), d), _fp(c), c.shape[0], N, L, d,
                          _ptr(out), L), (_stream_ptr(dev),))
        return out if out_dtype == torch.float32 else out.to(out_dtype)
    if not torch.cuda.is_available():
        raise RuntimeError("filter: no CUDA device; this package has no CPU path")
    dev = torch.device(device if device is not None else "cuda")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    src_h = src.detach().to(torch.float32).contiguous()
    ref_h = ref.detach().to(device="cpu", dtype=torch.float32).contiguous()
    out = torch.empty((N, L), dtype=torch.float32, pin_memory=True)
    if N == 0 or L == 0:
        return out.to(out_dtype)
    with torch.cuda.device(dev):
        ws = _filter_call(lib.sgp_filter_host, lib.sgp_filter_host_workspace_bytes, dev, N, d, L, order,
                          (C.c_void_p(src_h.data_ptr()), L, C.c_void_p(ref_h.data_ptr()), d, _fp(c), c.shape[0], N, L, d,
                           C.c_void_p(out.data_ptr()), L), (_stream_ptr(dev),))
        del ws   # sgp_filter_host has synchronised: nothing is in flight on the workspace
    return out if out_dtype == torch.float32 else out.to(out_dtype)


_side_streams = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One extra stream per device for build work that runs beside the main stream (batch packing of the blur groups)."""
    s = _side_streams.get(dev.index)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _side_streams[dev.index] = s
    return s
