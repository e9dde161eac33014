"""ctypes binding of ``include/sgp_lattice.h`` (the C ABI in ``libsgp_lattice.so``).

The library is compiled in-tree by ``csrc/build.py`` (plain nvcc, sm_100a).  There is no
CPU fallback: if the shared library is missing or fails to load, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libsgp_lattice.so")

SGP_OK = 0
SGP_ENOMEM = -6
SGP_SPLAT_AUTO, SGP_SPLAT_ATOMIC, SGP_SPLAT_GATHER = 0, 1, 2
MODE_AUTO, MODE_ATOMIC, MODE_GATHER, MODE_TILES, MODE_ROWS = 0, 1, 2, 3, 4   # Lattice.mvm(mode=...)
SGP_MAX_DIM = 126
SGP_MAX_ORDER = 7

# every symbol include/sgp_lattice.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sgp_abi_version", "sgp_last_error", "sgp_stencil_variance", "sgp_scale_factors", "sgp_slice_divisor",
    "sgp_build_points", "sgp_hash_capacity", "sgp_hash_insert", "sgp_number_workspace_bytes",
    "sgp_count_points", "sgp_number_points", "sgp_hash_seed", "sgp_hash_extend", "sgp_count_extension",
    "sgp_number_extension", "sgp_build_neighbours", "sgp_splat", "sgp_blur", "sgp_slice", "sgp_mvm", "sgp_debug_division_mismatches",
    "sgp_tiles_workspace_bytes", "sgp_tiles_prepare", "sgp_tiles_finalize", "sgp_splat_tiles", "sgp_slice_tiles",
    "sgp_grad_channels", "sgp_grad_pack", "sgp_grad_contract",
    "sgp_group_workspace_bytes", "sgp_group_prepare", "sgp_group_max_batches", "sgp_group_finalize",
    "sgp_group_prepare_async", "sgp_group_finalize_async",
    "sgp_remap_replay",
    "sgp_blur_groups_channel_block", "sgp_blur_groups", "sgp_mvm_rows_groups", "sgp_mvm_rows_groups_ex", "sgp_mvm_stage_splat_prezeroed", "sgp_sort_points_workspace_bytes", "sgp_sort_points",
    "sgp_permute_replay", "sgp_permute_replay_padded", "sgp_rowsort_workspace_bytes", "sgp_rowsort_padded", "sgp_build_rowsorted",
    "sgp_splat_rows", "sgp_cg_scratch_floats", "sgp_cg_apply", "sgp_cg_update", "sgp_cg_update_ex", "sgp_cg_direction", "sgp_cg_update_r", "sgp_cg_direction_x", "sgp_mvm_rows_groups_cg", "sgp_slice_ring_cg", "sgp_slice_ring_cg_supported", "sgp_cg_iteration", "sgp_pad_columns",
    "sgp_ring_enabled", "sgp_ring_splat_enabled", "sgp_ring_slice_enabled", "sgp_splat_ring_supported", "sgp_slice_ring_supported", "sgp_splat_rows_ring", "sgp_slice_ring",
    "sgp_hash_append_keys", "sgp_count_appended", "sgp_number_appended",
    "sgp_filter_workspace_bytes", "sgp_filter_host_workspace_bytes", "sgp_filter", "sgp_filter_host",
]


class LatticeView(C.Structure):
    """Mirror of ``struct sgp_lattice_view``."""

    _fields_ = [
        ("N", C.c_int64),
        ("M", C.c_int64),
        ("d", C.c_int32),
        ("order", C.c_int32),
        ("replay", C.c_void_p),
        ("nbr", C.c_void_p),
        ("csr_ptr", C.c_void_p),
        ("csr_ent", C.c_void_p),
        ("perm", C.c_void_p),
        ("fast", C.c_int32),
        ("replay_transposed", C.c_int32),
        ("replay_stride", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class TilesView(C.Structure):
    """Mirror of ``struct sgp_tiles_view``."""

    _fields_ = [
        ("N", C.c_int64),
        ("M", C.c_int64),
        ("S", C.c_int64),
        ("P", C.c_int64),
        ("d", C.c_int32),
        ("tile_points", C.c_int32),
        ("dict_cap", C.c_int32),
        ("reserved", C.c_int32),
        ("perm", C.c_void_p),
        ("tile_seg_ptr", C.c_void_p),
        ("seg_row", C.c_void_p),
        ("lidx", C.c_void_p),
        ("tile_w", C.c_void_p),
        ("seg_ent", C.c_void_p),
        ("tile_piece_ptr", C.c_void_p),
        ("piece_ptr", C.c_void_p),
        ("piece_row", C.c_void_p),
    ]


class BlurGroup(C.Structure):
    """Mirror of ``struct sgp_blur_group``."""

    _fields_ = [
        ("j0", C.c_int32),
        ("j1", C.c_int32),
        ("rows_cap", C.c_int32),
        ("zero_row", C.c_int32),
        ("n_batches", C.c_int64),
        ("batch_begin", C.c_void_p),
        ("src", C.c_void_p),
        ("lnb", C.c_void_p),
    ]


class SgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sgp_lattice error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once). Raises if it has not been built: no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python simplex-gp_b200/csrc/build.py` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
    fp = C.POINTER(C.c_float)
    L.sgp_abi_version.restype = i32
    L.sgp_abi_version.argtypes = []
    L.sgp_last_error.restype = C.c_char_p
    L.sgp_last_error.argtypes = []
    L.sgp_stencil_variance.restype = i32
    L.sgp_stencil_variance.argtypes = [fp, i32, fp]
    L.sgp_scale_factors.restype = i32
    L.sgp_scale_factors.argtypes = [i32, C.c_float, fp]
    L.sgp_slice_divisor.restype = C.c_float
    L.sgp_slice_divisor.argtypes = [i32]
    L.sgp_build_points.restype = i32
    L.sgp_build_points.argtypes = [vp, i64, i32, i64, fp, vp, vp, vp, vp, vp]
    L.sgp_hash_capacity.restype = i64
    L.sgp_hash_capacity.argtypes = [i64]
    L.sgp_hash_insert.restype = i32
    L.sgp_hash_insert.argtypes = [vp, vp, i64, i32, vp, i64, vp, vp, vp]
    L.sgp_number_workspace_bytes.restype = sz
    L.sgp_number_workspace_bytes.argtypes = [i64, i32]
    L.sgp_count_points.restype = i32
    L.sgp_count_points.argtypes = [vp, i64, vp, i64, i32, vp, sz, vp, C.POINTER(i64), C.POINTER(C.c_int32), vp]
    L.sgp_number_points.restype = i32
    L.sgp_number_points.argtypes = [vp, i64, vp, vp, vp, i64, i32, vp, i64, vp, vp, vp]
    L.sgp_hash_seed.restype = i32
    L.sgp_hash_seed.argtypes = [vp, i64, i32, vp, i64, vp, vp]
    L.sgp_hash_extend.restype = i32
    L.sgp_hash_extend.argtypes = [vp, vp, i64, i32, vp, i64, vp, i64, vp, vp, vp]
    L.sgp_count_extension.restype = i32
    L.sgp_count_extension.argtypes = [vp, i64, vp, i64, i32, vp, sz, vp, C.POINTER(i64), C.POINTER(C.c_int32), vp]
    L.sgp_number_extension.restype = i32
    L.sgp_number_extension.argtypes = [vp, i64, vp, vp, vp, i64, i32, vp, i64, i64, vp, vp, vp]
    L.sgp_hash_append_keys.restype = i32
    L.sgp_hash_append_keys.argtypes = [vp, i64, i32, vp, i64, vp, i64, vp, vp, vp]
    L.sgp_count_appended.restype = i32
    L.sgp_count_appended.argtypes = [vp, i64, vp, i64, vp, sz, vp, C.POINTER(i64), C.POINTER(C.c_int32), vp]
    L.sgp_number_appended.restype = i32
    L.sgp_number_appended.argtypes = [vp, i64, vp, vp, i64, i32, vp, i64, i64, vp, vp, vp]
    L.sgp_build_neighbours.restype = i32
    L.sgp_build_neighbours.argtypes = [vp, i64, i32, i32, vp, i64, vp, vp]
    pv = C.POINTER(LatticeView)
    L.sgp_splat.restype = i32
    L.sgp_splat.argtypes = [pv, vp, i64, i32, vp, i32, vp]
    L.sgp_blur.restype = i32
    L.sgp_blur.argtypes = [pv, fp, i32, i32, vp, vp, C.POINTER(C.c_int), vp]
    L.sgp_slice.restype = i32
    L.sgp_slice.argtypes = [pv, vp, i32, vp, i64, i32, vp]
    L.sgp_mvm.restype = i32
    L.sgp_mvm.argtypes = [pv, vp, i64, i32, fp, i32, vp, i64, vp, vp, i32, vp]
    pt = C.POINTER(TilesView)
    L.sgp_tiles_workspace_bytes.restype = sz
    L.sgp_tiles_workspace_bytes.argtypes = [i64, i32]
    L.sgp_tiles_prepare.restype = i32
    L.sgp_tiles_prepare.argtypes = [vp, i64, i32, i64, i32, vp, vp, sz, C.POINTER(i64), vp]
    L.sgp_tiles_finalize.restype = i32
    L.sgp_tiles_finalize.argtypes = [vp, vp, i64, i32, i32, i64, vp, sz, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32), vp]
    L.sgp_splat_tiles.restype = i32
    L.sgp_splat_tiles.argtypes = [pt, vp, i64, i32, vp, vp]
    L.sgp_slice_tiles.restype = i32
    L.sgp_slice_tiles.argtypes = [pt, vp, i32, vp, i64, i32, vp]
    L.sgp_grad_channels.restype = i32
    L.sgp_grad_channels.argtypes = [i32, i32]
    L.sgp_grad_pack.restype = i32
    L.sgp_grad_pack.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, vp, i64, vp]
    L.sgp_grad_contract.restype = i32
    L.sgp_grad_contract.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, i32, i32, vp, i64, vp, i64, vp]
    L.sgp_group_workspace_bytes.restype = sz
    L.sgp_group_workspace_bytes.argtypes = [i64]
    L.sgp_group_prepare.restype = i32
    L.sgp_group_prepare.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, vp, sz, C.POINTER(i64), vp]
    L.sgp_group_finalize.restype = i32
    L.sgp_group_finalize.argtypes = [vp, vp, i32, vp, i64, i64, i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, sz,
                                     C.POINTER(i64), C.POINTER(C.c_int32), vp]
    L.sgp_group_prepare_async.restype = i32
    L.sgp_group_prepare_async.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, vp, sz, vp, vp]
    L.sgp_group_finalize_async.restype = i32
    L.sgp_group_finalize_async.argtypes = [vp, vp, i32, vp, i64, i64, i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp]
    L.sgp_group_max_batches.restype = i64
    L.sgp_group_max_batches.argtypes = [i64, i64, i64]
    L.sgp_remap_replay.restype = i32
    L.sgp_remap_replay.argtypes = [vp, i64, vp, vp, vp]
    L.sgp_blur_groups_channel_block.restype = i32
    L.sgp_blur_groups_channel_block.argtypes = [i32]
    L.sgp_blur_groups.restype = i32
    L.sgp_blur_groups.argtypes = [C.POINTER(BlurGroup), i32, i64, i32, fp, i32, i32, vp, vp, C.POINTER(C.c_int), i32, vp]
    L.sgp_mvm_rows_groups.restype = i32
    L.sgp_mvm_rows_groups.argtypes = [pv, vp, vp, i64, C.POINTER(BlurGroup), i32, vp, i64, i32, fp, i32, vp, i64, vp, vp, i32, vp]
    L.sgp_mvm_stage_splat_prezeroed.restype = i32
    L.sgp_mvm_stage_splat_prezeroed.argtypes = [vp, vp, i64, i64, i64, vp, i64, i32, vp, i32, vp]
    L.sgp_mvm_rows_groups_ex.restype = i32
    L.sgp_mvm_rows_groups_ex.argtypes = [pv, vp, vp, i64, C.POINTER(BlurGroup), i32, vp, i64, i32, fp, i32, vp, i64, vp, vp, i32,
                                         i32, vp]
    L.sgp_mvm_rows_groups_cg.restype = i32
    L.sgp_mvm_rows_groups_cg.argtypes = [pv, vp, vp, i64, C.POINTER(BlurGroup), i32, vp, i64, i32, fp, i32, vp, i64, vp, vp, i32,
                                         i32, vp, vp, vp, vp, vp]
    L.sgp_pad_columns.restype = i32
    L.sgp_pad_columns.argtypes = [vp, i64, i32, vp, i64, i32, i64, vp]
    L.sgp_cg_iteration.restype = i32
    L.sgp_cg_iteration.argtypes = [pv, vp, vp, i64, C.POINTER(BlurGroup), i32, fp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp,
                                   vp, vp, C.c_float, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]
    L.sgp_slice_ring_cg_supported.restype = i32
    L.sgp_slice_ring_cg_supported.argtypes = [pv, vp, i32, vp, i64, i32, vp, i64]
    L.sgp_slice_ring_cg.restype = i32
    L.sgp_slice_ring_cg.argtypes = [pv, vp, i32, vp, i64, i32, vp, i64, vp, vp, vp, vp, vp]
    L.sgp_sort_points_workspace_bytes.restype = sz
    L.sgp_sort_points_workspace_bytes.argtypes = [i64]
    L.sgp_sort_points.restype = i32
    L.sgp_sort_points.argtypes = [vp, i64, i32, vp, vp, sz, vp]
    L.sgp_permute_replay.restype = i32
    L.sgp_permute_replay.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    L.sgp_permute_replay_padded.restype = i32
    L.sgp_permute_replay_padded.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    L.sgp_rowsort_workspace_bytes.restype = sz
    L.sgp_rowsort_workspace_bytes.argtypes = [i64, i32, i64]
    L.sgp_rowsort_padded.restype = i64
    L.sgp_rowsort_padded.argtypes = [i64, i32, i64]
    L.sgp_build_rowsorted.restype = i32
    L.sgp_build_rowsorted.argtypes = [vp, i64, i32, i64, i64, vp, vp, vp, vp, sz, vp]
    L.sgp_cg_scratch_floats.restype = sz
    L.sgp_cg_scratch_floats.argtypes = [i32]
    L.sgp_cg_apply.restype = i32
    L.sgp_cg_apply.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
    L.sgp_cg_update.restype = i32
    L.sgp_cg_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_float, i64, i32, vp, vp, vp, vp, vp]
    L.sgp_cg_update_ex.restype = i32
    L.sgp_cg_update_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_float, i32, i64, i32, vp, vp, vp, vp, vp]
    L.sgp_cg_direction.restype = i32
    L.sgp_cg_direction.argtypes = [vp, vp, vp, i64, i32, vp]
    L.sgp_cg_update_r.restype = i32
    L.sgp_cg_update_r.argtypes = [vp, vp, vp, vp, vp, C.c_float, i32, i64, i32, vp, vp, vp, vp, vp]
    L.sgp_cg_direction_x.restype = i32
    L.sgp_cg_direction_x.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    L.sgp_splat_rows.restype = i32
    L.sgp_splat_rows.argtypes = [vp, vp, i64, i64, i64, vp, i64, i32, vp, i32, vp]
    for fn in (L.sgp_ring_enabled, L.sgp_ring_splat_enabled, L.sgp_ring_slice_enabled):
        fn.restype = i32
        fn.argtypes = []
    L.sgp_splat_ring_supported.restype = i32
    L.sgp_splat_ring_supported.argtypes = [vp, i32]
    L.sgp_slice_ring_supported.restype = i32
    L.sgp_slice_ring_supported.argtypes = [pv, vp, i32]
    L.sgp_splat_rows_ring.restype = i32
    L.sgp_splat_rows_ring.argtypes = [vp, vp, i64, i64, i64, vp, i64, i32, vp, i32, vp]
    L.sgp_slice_ring.restype = i32
    L.sgp_slice_ring.argtypes = [pv, vp, i32, vp, i64, i32, vp]
    for fn in (L.sgp_filter_workspace_bytes, L.sgp_filter_host_workspace_bytes):
        fn.restype = sz
        fn.argtypes = [i64, i32, i32, i32, i64]
    for fn in (L.sgp_filter, L.sgp_filter_host):
        fn.restype = i32
        fn.argtypes = [vp, i64, vp, i64, fp, i32, i64, i32, i32, vp, i64, vp, sz, i64, C.POINTER(i64), vp]
    L.sgp_debug_division_mismatches.restype = i32
    L.sgp_debug_division_mismatches.argtypes = [i32, C.c_uint32, C.c_uint32, vp, vp]
    if L.sgp_abi_version() != 5:
        raise RuntimeError(f"{LIB_PATH}: ABI version {L.sgp_abi_version()} != 5, rebuild the library")
    _lib = L
    return L


def check(code: int) -> None:
    if code != SGP_OK:
        raise SgpError(code, lib().sgp_last_error().decode("utf-8", "replace"))
