/*
 * sgp_lattice.h -- C ABI of the B200-native Simplex-GP lattice filter.
 *
 * This is the drop-in boundary for the reference's native operator
 *
 *     filter(src[N,L], ref[N,d], coeffs[2r+1]) -> out[N,L]
 *
 * (pybind11 function exported by gpytorch_lattice_kernel/cpp/lattice.cpp:6-16 and
 * gpytorch_lattice_kernel/cuda/permutohedral_cuda.cpp:12-22 of the reference, called
 * from gpytorch_lattice_kernel/bilateral_kernel.py:95,111,119).  The reference rebuilds
 * the lattice inside every call.  Two forms are exported:
 *   - sgp_filter / sgp_filter_host (end of this file): the one-call form, same semantics
 *     as the reference's filter -- build, one product, nothing kept;
 *   - the stage entry points, so that a lattice is built once per hyper-parameter step and
 *     reused by every MVM of a solve (what sgp_filter itself is made of).
 *
 * Conventions
 *   - plain C, no torch types: raw pointers, sizes, and a cudaStream_t passed as void*.
 *   - every pointer documented "device" is device memory owned by the caller (PyTorch's
 *     caching allocator in the Python host); the library never allocates device memory:
 *     entry points that need scratch take a caller-provided workspace and have a
 *     *_workspace_bytes companion.
 *   - every function returns an int status: 0 = OK, negative = error; the message of
 *     the last error on the calling thread is available from sgp_last_error().
 *     Nothing calls exit() or throws across the boundary (the reference's CUDA path
 *     exits the process on a CUDA error, permutohedral_cuda_kernel.cu:24-32).
 *   - all launches go to the given stream and are asynchronous.  The entry points that
 *     synchronise it say so: the ones that return a count the host must size arrays with
 *     (sgp_count_points, sgp_count_extension, sgp_group_prepare, sgp_group_finalize,
 *     sgp_tiles_prepare, sgp_tiles_finalize), sgp_filter (through sgp_count_points) and
 *     sgp_filter_host (which returns with the result in host memory).
 *   - fp32 values (the reference CPU path is fp32-only, permutohedral.h:12,17),
 *     int16 keys, int8 ranks, int32 lattice indices.
 */
#ifndef SGP_LATTICE_H
#define SGP_LATTICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGP_ABI_VERSION 5

#define SGP_OK 0
#define SGP_EINVAL (-1)       /* bad argument */
#define SGP_ERANGE (-2)       /* a lattice coordinate does not fit int16 (reference keys are short) */
#define SGP_EOVERFLOW (-3)    /* hash table full or an index exceeds 32 bits */
#define SGP_ECUDA (-4)        /* CUDA runtime error */
#define SGP_EUNSUPPORTED (-5) /* shape outside what the kernels are built for */
#define SGP_ENOMEM (-6)       /* the caller's workspace was sized for fewer lattice points than the lattice has */

#define SGP_MAX_DIM 126   /* rank is a signed char in the reference (permutohedral.h:354) */
#define SGP_MAX_ORDER 7

/* status bits the build kernels OR into *status_flags (device int32) */
#define SGP_FLAG_KEY_RANGE 1
#define SGP_FLAG_TABLE_FULL 2

typedef void *sgp_stream_t; /* cudaStream_t */

int sgp_abi_version(void);
const char *sgp_last_error(void);

/* ---- host-side constants -------------------------------------------------------- */

/* second central moment of the stencil, fp32 sequential (permutohedral.h:203-219) */
int sgp_stencil_variance(const float *coeffs, int k, float *var_out);
/* scale_factor[i], i<d (permutohedral.h:372-390) */
int sgp_scale_factors(int d, float var, float *scale_out);
/* 1 + 2^-d in fp32, the divisor in slice (permutohedral.h:507) */
float sgp_slice_divisor(int d);

/* ---- stage 1: lattice construction ---------------------------------------------- */

/* Per-point geometry (permutohedral.h:397-465): elevate, nearest remainder-0 point,
 * rank sort, barycentric weights.
 *   x        device [N, ldx] fp32, first d columns used (x is already divided by the lengthscale)
 *   scale    HOST   [d] from sgp_scale_factors
 *   greedy   device [N, d+1] int16      rank  device [N, d+1] int8
 *   replay   device [N, d+1, 2] int32: word 1 receives the weight bits; word 0 (lattice
 *            index) is written by sgp_number_points
 *   status_flags device int32, OR-ed with SGP_FLAG_* */
int sgp_build_points(const float *x, int64_t N, int d, int64_t ldx, const float *scale,
                     int16_t *greedy, int8_t *rank, int32_t *replay, int32_t *status_flags,
                     sgp_stream_t stream);

/* number of 64-bit slots to allocate for n_keys insertions (power of two, load <= 1/2) */
int64_t sgp_hash_capacity(int64_t n_keys);

/* Lock-free find-or-create of the N*(d+1) vertex keys (permutohedral.h:467-474, table
 * semantics :58-94).  table: device [capacity] uint64, must be filled with 0xFF bytes.
 * slot_of: device [N*(d+1)] uint32 receives the slot of each point-vertex.
 * After the call every occupied slot holds the SMALLEST point-vertex id n*(d+1)+rem that
 * carries its key, which is what makes the numbering below deterministic. */
int sgp_hash_insert(const int16_t *greedy, const int8_t *rank, int64_t N, int d,
                    uint64_t *table, int64_t capacity, uint32_t *slot_of,
                    int32_t *status_flags, sgp_stream_t stream);

/* bytes of device scratch needed by sgp_count_points / sgp_number_points */
size_t sgp_number_workspace_bytes(int64_t N, int d);

/* First-touch numbering, part 1: marks the point-vertices that create a lattice point
 * (the owners left by sgp_hash_insert), prefix-sums the marks in point-vertex order and
 * returns M and the status flags.  Synchronises the stream.  Lattice point i is the i-th
 * key the reference's sequential scan over (n, rem) would have created (:73-79). */
int sgp_count_points(const uint64_t *table, int64_t capacity, const uint32_t *slot_of,
                     int64_t N, int d, void *workspace, size_t workspace_bytes,
                     const int32_t *status_flags, int64_t *M_out, int32_t *flags_out,
                     sgp_stream_t stream);

/* First-touch numbering, part 2: writes replay[...,0] (lattice index per point-vertex),
 * keys [M, d] int16 in first-touch order, and rewrites the table so that each occupied
 * slot maps key -> lattice index (used by sgp_build_neighbours). */
int sgp_number_points(uint64_t *table, int64_t capacity, const uint32_t *slot_of,
                      const int16_t *greedy, const int8_t *rank, int64_t N, int d,
                      const void *workspace, int64_t M, int32_t *replay, int16_t *keys,
                      sgp_stream_t stream);

/* ---- Extending a built lattice with more points ------------------------------------
 * The rectangular operator K(Xin, Xout) of the reference filters on the lattice of cat([Xout, Xin])
 * (gpytorch_lattice_kernel/bilateral_kernel.py:142-160) and rebuilds that lattice in every product.  First-touch
 * numbering is sequential over the points (permutohedral.h:73-79, 467-485), so the union lattice numbers the keys of
 * Xout exactly as the lattice of Xout alone and appends the keys only Xin touches: these calls add the new points to
 * an existing lattice without revisiting the old ones.
 *
 * sgp_hash_seed      fills an empty table (all 0xFF) with key -> index for keys[M_old]  (cost ~ M_old).
 * sgp_hash_extend    inserts the N_new*(d+1) new point-vertices (greedy_new / rank_new from sgp_build_points on the
 *                    new points) against it; slot_of_new: device [N_new*(d+1)].
 * sgp_count_extension  marks + scans the new point-vertices; *M_add_out = lattice points they create.  Synchronises.
 *                    Workspace: sgp_number_workspace_bytes(N_new, d).
 * sgp_number_extension  writes replay_new[..., 0] (indices into the union lattice), keys rows [M_old, M_old + M_add)
 *                    (keys: device [M_old + M_add, d] whose first M_old rows the caller has filled), and leaves the
 *                    table mapping every key of the union to its index (for sgp_build_neighbours / sgp_group_finalize).
 * Limits: N_new*(d+1) and M_old + N_new*(d+1) below 2^31. */
int sgp_hash_seed(const int16_t *keys, int64_t M_old, int d, uint64_t *table, int64_t capacity,
                  int32_t *status_flags, sgp_stream_t stream);
int sgp_hash_extend(const int16_t *greedy_new, const int8_t *rank_new, int64_t N_new, int d,
                    const int16_t *keys, int64_t M_old, uint64_t *table, int64_t capacity,
                    uint32_t *slot_of_new, int32_t *status_flags, sgp_stream_t stream);
int sgp_count_extension(const uint64_t *table, int64_t capacity, const uint32_t *slot_of_new, int64_t N_new,
                        int d, void *workspace, size_t workspace_bytes, const int32_t *status_flags,
                        int64_t *M_add_out, int32_t *flags_out, sgp_stream_t stream);
int sgp_number_extension(uint64_t *table, int64_t capacity, const uint32_t *slot_of_new,
                         const int16_t *greedy_new, const int8_t *rank_new, int64_t N_new, int d,
                         const void *workspace, int64_t M_old, int64_t M_add, int32_t *replay_new,
                         int16_t *keys, sgp_stream_t stream);

/* ---- Merging key lists: the numbering step of a point-sharded build ------------------------
 * Under point sharding rank g builds the lattice of its own points [lo_g, hi_g) and gets its keys in local first-touch
 * order.  Ranks own contiguous point ranges, so the reference's sequential numbering (permutohedral.h:73-79,467-485) of
 * the whole point set is: rank 0's keys, then rank 1's keys that rank 0 does not hold, ... each list in its own order.
 * Every rank replays that merge over the all-gathered lists, with the same result everywhere:
 *   table seeded with the keys so far (sgp_hash_seed, or empty = all 0xFF for the first list; capacity >=
 *   sgp_hash_capacity(total keys)), then per list
 * sgp_hash_append_keys   find-or-insert the m_new DISTINCT keys new_keys[m_new, d]; slot_of_new: device [m_new].
 * sgp_count_appended     *M_add_out = keys of the list that were not in the table.  Synchronises.  Workspace:
 *                        sgp_number_workspace_bytes(ceil(m_new / (d+1)), d) bytes or more (4 m_new + scan scratch).
 * sgp_number_appended    keys rows [M_old, M_old + M_add) <- the new keys in list order; map_out (device [m_new], may be
 *                        NULL) <- index of every listed key in the merged numbering; the table ends mapping every key of
 *                        the union to its index (for sgp_build_neighbours / sgp_group_finalize). */
int sgp_hash_append_keys(const int16_t *new_keys, int64_t m_new, int d, const int16_t *keys, int64_t M_old,
                         uint64_t *table, int64_t capacity, uint32_t *slot_of_new, int32_t *status_flags,
                         sgp_stream_t stream);
int sgp_count_appended(const uint64_t *table, int64_t capacity, const uint32_t *slot_of_new, int64_t m_new,
                       void *workspace, size_t workspace_bytes, const int32_t *status_flags, int64_t *M_add_out,
                       int32_t *flags_out, sgp_stream_t stream);
int sgp_number_appended(uint64_t *table, int64_t capacity, const uint32_t *slot_of_new, const int16_t *new_keys,
                        int64_t m_new, int d, const void *workspace, int64_t M_old, int64_t M_add, int32_t *map_out,
                        int16_t *keys, sgp_stream_t stream);

/* Neighbour table of the blur (permutohedral.h:539-545): nbr[j, i, t] = lattice index of
 * the key "key[i] - o on every stored coordinate, then key[i][j] + o*d on coordinate j when
 * j < d" (axis j = d only shifts), t enumerating o = -r..-1, 1..r; -1 if absent.
 * int16 wrap-around as in the reference.  nbr: device [(d+1), M, 2r] int32. */
int sgp_build_neighbours(const int16_t *keys, int64_t M, int d, int order,
                         const uint64_t *table, int64_t capacity, int32_t *nbr,
                         sgp_stream_t stream);

/* ---- stages 2-4: the MVM on a built lattice -------------------------------------- */

typedef struct sgp_lattice_view {
    int64_t N;               /* points */
    int64_t M;               /* lattice points */
    int32_t d;               /* input dimension */
    int32_t order;           /* stencil half-width r */
    const int32_t *replay;   /* device [N, d+1, 2] {lattice index, weight bits} */
    const int32_t *nbr;      /* device [(d+1), M, 2r] */
    const uint32_t *csr_ptr; /* device [M+1] or NULL: row starts into csr_ent (ordered-gather splat) */
    const int32_t *csr_ent;  /* device or NULL: {point, weight bits} sorted by lattice row, point-vertex order inside a
                                row -- sgp_build_rowsorted's `ent`, in its interleaved storage order (csr_ptr holds
                                row-sorted positions; bit 31 of the point word is ignored) */
    const uint32_t *perm;    /* device [N] or NULL.  When set, row p of replay describes point perm[p]: splat and
                                slice walk the points in that (locality) order and address src / out rows through it */
    int32_t fast;            /* 0: the reference's arithmetic, one rounding per product and per sum (bit-exact on the
                                deterministic path); 1: fused multiply-adds and one division per output in slice
                                (differs by rounding only, ~1e-7 relative) */
    int32_t replay_transposed; /* 0: replay is [N, d+1, 2]; 1: [d+1, N, 2] (what sgp_permute_replay can produce: the
                                  points of a warp then read one contiguous run per vertex) */
    int32_t replay_stride;     /* entries per point of a non-transposed replay table: 0 or d+1 = dense; d+2 for the
                                  table sgp_permute_replay_padded makes when d+1 is odd (one {0, 0.0f} filler per point,
                                  so that every point starts 16-byte aligned and the slice reads entry PAIRS) */
    int32_t reserved_;
} sgp_lattice_view;

#define SGP_SPLAT_AUTO 0
#define SGP_SPLAT_ATOMIC 1 /* vectorised red.global.add scatter */
#define SGP_SPLAT_GATHER 2 /* CSR gather, deterministic, reference accumulation order */
#define SGP_SPLAT_ATOMIC_ACCUMULATE 5 /* as ATOMIC, but adds to `values` as they are (no memset): the view of a point
                                         range contributes its part of a splat that is fed chunk by chunk */

/* values[M, L] = sum_{point-vertices} weight * src[n, :]   (permutohedral.h:478-479) */
int sgp_splat(const sgp_lattice_view *lat, const float *src, int64_t lds, int L,
              float *values, int mode, sgp_stream_t stream);
/* d+1 stencil passes (permutohedral.h:526-556), ping-pong between buf0 (input) and buf1.
 * coeffs: HOST [2r+1].  *result_in_buf1 tells where the blurred values ended. */
int sgp_blur(const sgp_lattice_view *lat, const float *coeffs, int k, int L,
             float *buf0, float *buf1, int *result_in_buf1, sgp_stream_t stream);
/* out[n, :L_out] = sum_rem (w * values[idx, :L_out]) / (1 + 2^-d)   (permutohedral.h:497-510).  values: [M, L];
 * L_out <= L lets the lattice rows be padded to a multiple of 4 channels (16-byte vectors) while out keeps the
 * caller's width and alignment (then written one channel at a time). */
int sgp_slice(const sgp_lattice_view *lat, const float *values, int L, float *out,
              int64_t ldo, int L_out, sgp_stream_t stream);
/* splat -> blur -> slice; buf0/buf1: device [M, L] fp32 scratch each */
int sgp_mvm(const sgp_lattice_view *lat, const float *src, int64_t lds, int L,
            const float *coeffs, int k, float *out, int64_t ldo,
            float *buf0, float *buf1, int splat_mode, sgp_stream_t stream);

/* ---- locality tiles: shared-memory staged splat and slice ------------------------------
 *
 * Points are sorted so that points sharing lattice vertices are adjacent and cut into tiles of
 * tile_points points.  Per tile, the distinct lattice rows it touches form its dictionary
 * (seg_row[tile_seg_ptr[t] .. tile_seg_ptr[t+1])) and its point-vertices are grouped by dictionary
 * entry into segments (seg_ptr / seg_ent).  See simplex-gp_b200/csrc/sgp_tiles.cu.  This is an
 * internal acceleration structure; the observable lattice (replay, keys, nbr) is unchanged. */
typedef struct sgp_tiles_view {
    int64_t N;                      /* points */
    int64_t M;                      /* lattice points */
    int64_t S;                      /* segments = sum of dictionary sizes */
    int64_t P;                      /* splat pieces (segments cut into runs of at most 8 entries) */
    int32_t d;
    int32_t tile_points;            /* T */
    int32_t dict_cap;               /* largest dictionary */
    int32_t reserved;
    const uint32_t *perm;           /* device [N]: sorted position -> point (sgp_sort_points) */
    const uint32_t *tile_seg_ptr;   /* device [n_tiles+1]                                    (slice) */
    const int32_t *seg_row;         /* device [S] lattice row of each dictionary entry, in the order in which the
                                       stage addresses the lattice values                     (slice) */
    const uint16_t *lidx;           /* device [N*(d+1)] dictionary index of each sorted point-vertex (slice) */
    const float *tile_w;            /* device [N*(d+1)] weight of each sorted point-vertex    (slice) */
    const int32_t *seg_ent;         /* device [N*(d+1), 2] {point index inside its tile, weight bits}, grouped by
                                       segment                                                (splat) */
    const uint32_t *tile_piece_ptr; /* device [n_tiles+1]                                     (splat) */
    const uint32_t *piece_ptr;      /* device [P+1] entry range of each piece                 (splat) */
    const int32_t *piece_row;       /* device [P] lattice row of each piece                   (splat) */
} sgp_tiles_view;

size_t sgp_tiles_workspace_bytes(int64_t N, int d);
/* segment count of the tiling of the points in the order perm (device [N], from sgp_sort_points);
 * synchronises the stream; returns S */
int sgp_tiles_prepare(const int32_t *replay, int64_t N, int d, int64_t M, int tile_points,
                      const uint32_t *perm, void *workspace, size_t workspace_bytes, int64_t *S_out,
                      sgp_stream_t stream);
/* fills the tile arrays (sizes as in sgp_tiles_view) from the workspace left by sgp_tiles_prepare;
 * synchronises the stream; *max_dict_out = largest dictionary */
int sgp_tiles_finalize(const int32_t *replay, const uint32_t *perm, int64_t N, int d, int tile_points,
                       int64_t S, void *workspace, size_t workspace_bytes, uint32_t *seg_ptr,
                       int32_t *seg_row, int32_t *seg_ent, uint32_t *tile_seg_ptr, uint16_t *lidx,
                       float *tile_w, int32_t *max_dict_out, sgp_stream_t stream);
/* values[M, L] = splat(src): one vector reduction per piece (values is zeroed inside) */
int sgp_splat_tiles(const sgp_tiles_view *tiles, const float *src, int64_t lds, int L, float *values,
                    sgp_stream_t stream);
/* out = slice(values), dictionary rows staged in shared memory; fast as in sgp_lattice_view */
int sgp_slice_tiles(const sgp_tiles_view *tiles, const float *values, int L, float *out, int64_t ldo,
                    int fast, sgp_stream_t stream);

/* ---- blur groups: several axes per launch, staged through shared memory ------------------
 *
 * For a range of consecutive axes [j0, j1) the lattice points split into classes that the passes of those
 * axes never leave; a CTA holding whole classes in shared memory runs all j1-j0 passes on chip (see
 * simplex-gp_b200/csrc/sgp_groups.cu).  A lattice's axes 0..d are covered by a chain of groups; stage g
 * gathers its input rows from stage g-1's output order (src) and writes its own order contiguously.
 * The arithmetic per pass is that of sgp_blur, so values are bit-identical to the per-axis path.  The slice
 * after the last stage reads through a replay table remapped to that stage's order (sgp_remap_replay). */
typedef struct sgp_blur_group {
    int32_t j0, j1;              /* axis range [j0, j1) */
    int32_t rows_cap;            /* largest number of rows of one CTA batch */
    int32_t zero_row;            /* 512 or 1024: the `cap` class the group was finalised with (cap <= 512 -> 512); it is
                                    the index absent neighbours carry in lnb and selects the kernel variant */
    int64_t n_batches;
    const uint32_t *batch_begin; /* device [n_batches+1] positions */
    const int32_t *src;          /* device [M] input row of every position */
    const uint16_t *lnb;         /* device [M*(j1-j0)*2r]: per batch [axis][row][t] batch-local neighbour positions;
                                    absent = 512 (batches of up to 512 rows) or 1024, the kernel's all-zero row */
} sgp_blur_group;

size_t sgp_group_workspace_bytes(int64_t M);
/* Sort the lattice points by class of the axis range: order[p] (device [M]) = lattice index at position p,
 * pos = its inverse, class_start[p] = first position of p's class.  Synchronises; *max_class_out = largest class. */
int sgp_group_prepare(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                      uint32_t *class_start, void *workspace, size_t workspace_bytes, int64_t *max_class_out,
                      sgp_stream_t stream);
/* Pack whole classes greedily into CTA batches of at most `cap` rows and fill the group tables.
 * batch_begin: device [sgp_group_max_batches(M, cap, max_class) + 1].  prev_pos: pos of the previous stage
 * (NULL for the first stage, whose input is in lattice-index order).  Synchronises; *n_batches_out = batches
 * used, *max_rows_out = rows of the largest batch. */
int64_t sgp_group_max_batches(int64_t M, int64_t cap, int64_t max_class);
/* Neighbours: from nbr (sgp_build_neighbours) when non-NULL, else looked up in the key hash table as left by
 * sgp_number_points (keys, d, table, capacity) -- the whole-lattice neighbour table then need not exist. */
int sgp_group_finalize(const int32_t *nbr, const int16_t *keys, int d, const uint64_t *table, int64_t capacity,
                       int64_t M, int order, int j0, int j1, const uint32_t *order_of,
                       const uint32_t *pos, const uint32_t *class_start, const uint32_t *prev_pos,
                       int64_t cap, int64_t max_batches, uint32_t *batch_begin, int32_t *src, uint16_t *lnb,
                       void *workspace, size_t workspace_bytes, int64_t *n_batches_out, int32_t *max_rows_out,
                       sgp_stream_t stream);
/* The same launches without the read-back: nothing synchronises, the figures land in `result` (device uint32[8]:
 * [0] largest class, [1] batches, [2] rows of the largest batch, [3] != 0: the classes did not fit / max_batches too
 * small, [4] != 0: a neighbour fell outside its batch).  A caller that knows the axis ranges (from the last lattice of
 * the same shape) runs sgp_group_prepare_async for every range on one stream, sgp_group_finalize_async on a second one
 * behind an event, and reads all results with ONE synchronisation; if a result is bad it falls back to the synchronous
 * pair above.  sgp_group_prepare_async needs the workspace only until the next prepare may overwrite it;
 * sgp_group_finalize_async needs none. */
int sgp_group_prepare_async(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                            uint32_t *class_start, void *workspace, size_t workspace_bytes, uint32_t *result,
                            sgp_stream_t stream);
int sgp_group_finalize_async(const int32_t *nbr, const int16_t *keys, int d, const uint64_t *table, int64_t capacity,
                             int64_t M, int order, int j0, int j1, const uint32_t *order_of, const uint32_t *pos,
                             const uint32_t *class_start, const uint32_t *prev_pos, int64_t cap, int64_t max_batches,
                             uint32_t *batch_begin, int32_t *src, uint16_t *lnb, uint32_t *result, sgp_stream_t stream);
/* replay_out[q] = {pos[replay[q].index], replay[q].weight bits}, q < total */
int sgp_remap_replay(const int32_t *replay, int64_t total, const uint32_t *pos, int32_t *replay_out,
                     sgp_stream_t stream);
/* channels staged per CTA for L channels */
int sgp_blur_groups_channel_block(int L);
/* run the chain: buf0 (lattice-index order) -> ... ; *result_in_buf1 tells where the last stage wrote;
 * fast as in sgp_lattice_view */
int sgp_blur_groups(const sgp_blur_group *groups, int n_groups, int64_t M, int order, const float *coeffs,
                    int k, int L, float *buf0, float *buf1, int *result_in_buf1, int fast, sgp_stream_t stream);

/* The production chain in one call: sgp_splat_rows -> sgp_blur_groups -> sgp_slice.  slice_view->replay addresses
 * the lattice values in the order the last group stage leaves them (sgp_permute_replay with that stage's pos);
 * ent / seg_row / n_entries: the row-sorted entries of sgp_build_rowsorted (n_entries = sgp_rowsort_padded(...));
 * buf0 / buf1: device [M, Lv] scratch, Lv = L or L rounded up to a multiple of 4 (see sgp_slice). */
int sgp_mvm_rows_groups(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                        int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src, int64_t lds, int L,
                        const float *coeffs, int k, float *out, int64_t ldo, float *buf0, float *buf1, int Lv,
                        sgp_stream_t stream);

/* The same chain with the memset of the splat buffer taken off the critical path.  flags: SGP_MVM_PREZEROED -- buf0
 * holds zeros on entry (no memset in front of the splat); SGP_MVM_ZERO_AFTER -- buf0 is left zeroed on exit: with an
 * odd number of group stages it is zeroed on an internal side stream while the slice runs (a parallel branch when the
 * call is captured into a CUDA graph), with an even number after the slice.  A graph captured with both flags on
 * private buffers (Lattice.capture) replays without ever waiting for the memset.
 * SGP_MVM_SRC_PADDED -- src has L rounded up to a multiple of 4 columns (lds >= that; the extra columns zero) while out
 * has L: what a caller passes after
 * copying a ragged block (L = 11: the reference's training block [y | 10 probes]) into a zero-padded one, so that the
 * splat gathers 16-byte vectors (config A with 11 columns: 236 -> 200 us per MVM including the copy, sgp_pad_columns). */
/* The splat stage of that chain alone: sgp_splat_rows without its memset -- `values` must hold zeros on entry. */
int sgp_mvm_stage_splat_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                  const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream);
/* The zero-padded copy SGP_MVM_SRC_PADDED refers to: dst[n, 0..Lv) = src[n, 0..L), then zeros.  Lv % 4 == 0, dst 16-byte
 * aligned, ldd % 4 == 0. */
/* Lv may exceed L rounded up to a multiple of 4: 12 (or 9-11) columns on 16-channel lattice rows keep every 64-byte row
 * gather inside one 128-byte line (config A, 12 columns: 181 -> 172 us per MVM); splat and slice then leave the spare lanes idle. */
int sgp_pad_columns(const float *src, int64_t lds, int L, float *dst, int64_t ldd, int Lv, int64_t N, sgp_stream_t stream);
#define SGP_MVM_PREZEROED 1
#define SGP_MVM_ZERO_AFTER 2
#define SGP_MVM_SRC_PADDED 4
int sgp_mvm_rows_groups_ex(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                           int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src, int64_t lds,
                           int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0, float *buf1, int Lv,
                           int flags, sgp_stream_t stream);
/* One CG iteration's product AND the sweep that follows it (sgp_cg_apply below): out = s*K*src + noise*src,
 * pAp[l] = sum_n src[n,l]*out[n,l].  With the TMA-ring slice the sweep runs in the slice's epilogue (the point's row of
 * src is read there, s*K*src + noise*src is what gets stored, per-CTA partial dot products are summed by a one-block
 * second stage); otherwise the slice is followed by sgp_cg_apply.  Contiguous blocks only: lds = ldo = L, and Lv = L or
 * (L % 4 == 0) a multiple of 4 above it -- 12 columns on 16-channel lattice rows, see sgp_mvm_rows_groups_ex; s, noise:
 * device scalars; pAp: device [L]; scratch: device [sgp_cg_scratch_floats(L)]; flags: SGP_MVM_PREZEROED / ZERO_AFTER. */
int sgp_mvm_rows_groups_cg(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                           int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src, int64_t lds,
                           int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0, float *buf1, int Lv,
                           int flags, const float *s, const float *noise, float *pAp, float *scratch, sgp_stream_t stream);
/* The slice of that chain alone, and whether it applies to a shape (16-byte vectors, ring slice selected). */
int sgp_slice_ring_cg_supported(const sgp_lattice_view *lat, const float *values, int L, const float *out, int64_t ldo,
                                int L_out, const float *P, int64_t ldp);
int sgp_slice_ring_cg(const sgp_lattice_view *lat, const float *values, int L, float *out, int64_t ldo, int L_out,
                      const float *P, int64_t ldp, const float *s, const float *noise, float *pAp, float *scratch,
                      sgp_stream_t stream);

/* ---- stage 5: lengthscale-gradient pass (bilateral_kernel.py:97-124) -------------------
 *
 * The reference filters one N x 2L(1+d) block [g | g(x)x | v | v(x)x] with the derivative stencil and
 * contracts it into grad_x (:113-122).  Channels are independent under the filter, so they are grouped
 * per RHS column l -- block(l) = [g_l, g_l x_1..x_d, v_l, v_l x_1..x_d], 2(1+d) channels -- and
 * processed nl columns at a time: sgp_grad_pack -> sgp_mvm (derivative stencil, lattice built with
 * that stencil's variance) -> sgp_grad_contract.  See simplex-gp_b200/csrc/sgp_grad.cu. */
int sgp_grad_channels(int d, int nl); /* 2(1+d)*nl */
/* packed[N, ldp] <- blocks of columns [l0, l0+nl) of g[N, ldg], v[N, ldv] with x[N, ldx] (all device) */
int sgp_grad_pack(const float *g, int64_t ldg, const float *v, int64_t ldv, const float *x, int64_t ldx,
                  int64_t N, int d, int l0, int nl, float *packed, int64_t ldp, sgp_stream_t stream);
/* grad_x[N, ldgx] (+)= sum over the chunk's columns of the reference expression (:122); first != 0
 * starts the sum at zero, last != 0 applies the factor -2.  Chunks must be visited in increasing l0.
 * grad_src (device [N, ldgs], may be NULL) receives the filtered g columns (:123). */
int sgp_grad_contract(const float *filtered, int64_t ldp, const float *g, int64_t ldg, const float *v,
                      int64_t ldv, const float *x, int64_t ldx, int64_t N, int d, int l0, int nl,
                      int first, int last, float *grad_x, int64_t ldgx, float *grad_src, int64_t ldgs,
                      sgp_stream_t stream);

/* ---- batched conjugate gradients: the vector updates either side of the MVM ------------------------------
 * The reference leaves the solve of (s K + noise I) X = B to GPyTorch (experiments/train_simplexgp.py:29-48).  One
 * CG iteration on an [N, L] block of right-hand sides is one lattice MVM (KP = K P) followed by
 *   sgp_cg_apply      AP = s*KP + noise*P in place on KP;  pAp[l] = sum_n P*AP
 *   sgp_cg_update     alpha = rs/max(pAp,1e-30);  X += alpha*P;  R -= alpha*AP;  rs_new = sum_n R*R;
 *                     beta = rs_new/max(rs,1e-30);  rs <- rs_new;  *done = all_l sqrt(rs_new[l])/bnorm[l] < tol
 *   sgp_cg_direction  P = R + beta*P
 * alpha_out / beta_out: device [L], this iteration's coefficients (rows of the Lanczos tridiagonals).  Blocks are
 * device [N, L] fp32, row-major, no row padding; 1 <= L <= 256.  s, noise: device scalars.  scratch: device
 * [sgp_cg_scratch_floats(L)] floats.  The dot products are summed in a fixed order (deterministic). */
size_t sgp_cg_scratch_floats(int L);
int sgp_cg_apply(float *AP, const float *P, const float *s, const float *noise, int64_t N, int L,
                 float *pAp, float *scratch, sgp_stream_t stream);
int sgp_cg_update(float *X, float *R, const float *P, const float *AP, float *rs, const float *pAp,
                  const float *bnorm, float tol, int64_t N, int L, float *alpha_out, float *beta_out,
                  int32_t *done, float *scratch, sgp_stream_t stream);
/* ... with the stopping rule chosen: SGP_CG_ALL_COLUMNS as above; SGP_CG_MEAN: *done = mean over the columns with a
 * non-zero right-hand side (bnorm > 1e-29) of sqrt(rs_new[l])/bnorm[l] < tol -- the rule of GPyTorch's linear_cg
 * (residual_norm.mean() < tolerance), which the reference's cg_tolerance setting refers to
 * (experiments/train_simplexgp.py:34-37). */
#define SGP_CG_ALL_COLUMNS 0
#define SGP_CG_MEAN 1
int sgp_cg_update_ex(float *X, float *R, const float *P, const float *AP, float *rs, const float *pAp,
                     const float *bnorm, float tol, int criterion, int64_t N, int L, float *alpha_out,
                     float *beta_out, int32_t *done, float *scratch, sgp_stream_t stream);
int sgp_cg_direction(float *P, const float *R, const float *beta, int64_t N, int L, sgp_stream_t stream);
/* The same iteration with X += alpha*P moved from the update sweep into the direction sweep (which reads P anyway):
 *   sgp_cg_update_r     alpha = rs/max(pAp,1e-30);  R -= alpha*AP;  rs_new, beta, done as sgp_cg_update_ex
 *   sgp_cg_direction_x  X += alpha*P;  P = R + beta*P
 * eight passes over [N, L] per iteration instead of nine, the same arithmetic in the same order.  When the iteration
 * stops after an update the caller adds the last alpha*P to X itself. */
int sgp_cg_update_r(float *R, const float *AP, float *rs, const float *pAp, const float *bnorm, float tol,
                    int criterion, int64_t N, int L, float *alpha_out, float *beta_out, int32_t *done, float *scratch,
                    sgp_stream_t stream);
int sgp_cg_direction_x(float *P, const float *R, float *X, const float *alpha, const float *beta, int64_t N, int L,
                       sgp_stream_t stream);
/* One whole CG iteration on the production chain (row-sorted splat -> blur groups -> slice with the CG epilogue ->
 * sgp_cg_update_r -> sgp_cg_direction_x) enqueued by one call; iteration `it` writes alphas[it, :], betas[it, :],
 * done[it] and copies done[it] to done_host[it] (pinned host memory, may be NULL) on the stream.  Blocks are unpadded
 * [N, L] (L % 4 == 0 or L <= 4 for the lattice side); buf0 / buf1: [M, Lv] work buffers; flags as sgp_mvm_rows_groups_ex;
 * X is complete after every call (the deferred X += alpha P is part of it). */
int sgp_cg_iteration(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row, int64_t n_entries,
                     const sgp_blur_group *groups, int n_groups, const float *coeffs, int k, float *buf0, float *buf1,
                     int flags, float *X, float *R, float *P, float *AP, float *rs, float *pAp, const float *bnorm,
                     const float *s, const float *noise, float tol, int criterion, int L, int Lv, float *alphas,
                     float *betas, int32_t *done, int32_t *done_host, int it, float *scratch, sgp_stream_t stream);

/* ---- row-sorted splat ("segmented gather") -------------------------------------------------
 * The point-vertices sorted by lattice row, point-vertex order within a row (the reference's accumulation order),
 * padded with zero-weight entries to n_entries = sgp_rowsort_padded(N, d, fill_rows), a multiple of 64:
 *   ent      device [n_entries, 2] int32 {point | row-start flag in bit 31, weight bits}; the flag marks the first
 *            entry of a lattice row.  STORAGE ORDER: the splat consumes 8 entries per thread as four 16-byte pieces, and
 *            the entries are stored in groups of 64 with the pieces interleaved over the group's 8 segments -- entry
 *            e = 64 g + 8 s + 2 p + h lives at position 64 g + 16 p + 2 s + h -- so that a warp instruction reads one
 *            contiguous 128-byte run;
 *   seg_row  device [n_entries / 4] int32: lattice row of every fourth entry;
 *   ent_row  optional (may be NULL) device [n_entries] int32: lattice row of every entry.
 * The encoding needs every lattice row 0..M-1 to own an entry.  That holds for the lattice of the points themselves
 * (fill_rows = 0); for a subset of the points on the full key set (a rank's share under point sharding) pass
 * fill_rows = M: one weightless filler entry per row is sorted in.
 * sgp_splat_rows gives every thread 8 consecutive entries -- 8.5 bytes of index stream per point-vertex: balanced
 * whatever the row lengths, one vector reduction per run of equal rows (values is zeroed inside). */
size_t sgp_rowsort_workspace_bytes(int64_t N, int d, int64_t fill_rows);
int64_t sgp_rowsort_padded(int64_t N, int d, int64_t fill_rows);
int sgp_build_rowsorted(const int32_t *replay, int64_t N, int d, int64_t M, int64_t fill_rows, int32_t *ent,
                        int32_t *ent_row, int32_t *seg_row, void *workspace, size_t workspace_bytes,
                        sgp_stream_t stream);
/* src: [N, lds] with L_src columns; values: [M, L], L >= L_src (columns L_src..L-1 receive zeros): as for sgp_slice,
 * the lattice rows may be padded to a multiple of 4 channels */
int sgp_splat_rows(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                   const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream);

/* ---- splat and slice with TMA-prefetched index streams (the production kernels; simplex-gp_b200/csrc/sgp_ring.cu) ----
 * Same contracts as sgp_splat_rows / sgp_slice, which forward here unless the environment says SGP_RING=0 or the shape
 * is outside what a warp covers (more than 32 channel chunks, a permuted or transposed replay table, d > 63).
 * Persistent warps stream the row-sorted entries / the replay table into shared-memory rings with cp.async.bulk +
 * mbarrier; sgp_splat_rows_ring writes every lattice row that lies inside one warp tile with a plain store and zeroes
 * (then reduces into) only the rows that cross a tile boundary -- `values` is NOT memset. */
int sgp_ring_enabled(void);        /* SGP_RING (default 1) */
int sgp_ring_splat_enabled(void);  /* ... and SGP_RING_SPLAT (default 1; used for dense lattices only) */
int sgp_ring_slice_enabled(void);  /* ... and SGP_RING_SLICE (default 1) */
int sgp_splat_ring_supported(const float *values, int L);
int sgp_slice_ring_supported(const sgp_lattice_view *lat, const float *values, int L);
int sgp_splat_rows_ring(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                        const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream);
int sgp_slice_ring(const sgp_lattice_view *lat, const float *values, int L, float *out, int64_t ldo, int L_out,
                   sgp_stream_t stream);

/* ---- locality order of the points --------------------------------------------------------
 * perm (device [N]): the points in lexicographic order of their remainder-0 lattice point, so that points sharing
 * lattice vertices are adjacent.  A replay table re-ordered with sgp_permute_replay plus sgp_lattice_view.perm makes
 * splat and slice walk the points in that order (neighbouring threads then touch the same lattice rows). */
size_t sgp_sort_points_workspace_bytes(int64_t N);
int sgp_sort_points(const int16_t *greedy, int64_t N, int d, uint32_t *perm, void *workspace,
                    size_t workspace_bytes, sgp_stream_t stream);
/* replay_out[p, r] (or [r, p] when transposed != 0) = {pos ? pos[replay[perm ? perm[p] : p, r].index] : that index,
 * weight bits}; perm and pos may be NULL */
int sgp_permute_replay(const int32_t *replay, const uint32_t *perm, const uint32_t *pos, int64_t N, int d,
                       int transposed, int32_t *replay_out, sgp_stream_t stream);

/* ---- the reference operator in one call --------------------------------------------------------
 *
 *     out[N, L] = filter(src[N, L], ref[N, d], coeffs[2r+1])
 *
 * exactly what gpytorch_lattice_kernel/cpp/lattice.cpp:6-16 and cuda/permutohedral_cuda.cpp:12-22 export:
 * the lattice of `ref` (positions already divided by the lengthscale) is built, `src` is splatted, blurred
 * along the d+1 axes with the stencil `coeffs` (HOST pointer, k = 2r+1 values) and sliced; nothing is kept
 * (permutohedral.h:259-340).  Values are those of the stage entry points with the fused-multiply-add
 * arithmetic (1e-7 relative from the reference's).
 *
 *   sgp_filter       src / ref / out are DEVICE pointers with leading dimensions lds / ldx / ldo (elements);
 *                    asynchronous on `stream` apart from the one synchronisation that fetches the number of
 *                    lattice points M.
 *   sgp_filter_host  src / ref / out are HOST pointers (pinned memory is copied asynchronously, pageable memory
 *                    through the driver's staging); the uploads, the filter and the download run inside and the
 *                    call returns with `out` complete.  The RHS block is uploaded on a second stream while the
 *                    lattice is being built.
 *
 * workspace: device memory, 256-byte aligned, at least sgp_filter[_host]_workspace_bytes(N, d, L, r, M_max)
 * bytes.  M_max is the caller's bound on the number of lattice points the workspace must hold (M_max <= 0 or
 * M_max > N(d+1): the worst case N(d+1), e.g. 9e6 x (72 B neighbour table + 2 x 4L B values) at N = 1M, d = 8).
 * *M_out (may be NULL) receives the actual M; if it exceeds M_max the call fails with SGP_ENOMEM and may be
 * repeated with a workspace sized for *M_out. */
size_t sgp_filter_workspace_bytes(int64_t N, int d, int L, int order, int64_t M_max);
size_t sgp_filter_host_workspace_bytes(int64_t N, int d, int L, int order, int64_t M_max);
int sgp_filter(const float *src, int64_t lds, const float *ref, int64_t ldx, const float *coeffs, int k,
               int64_t N, int L, int d, float *out, int64_t ldo, void *workspace, size_t workspace_bytes,
               int64_t M_max, int64_t *M_out, sgp_stream_t stream);
int sgp_filter_host(const float *src_host, int64_t lds, const float *ref_host, int64_t ldx, const float *coeffs,
                    int k, int64_t N, int L, int d, float *out_host, int64_t ldo, void *workspace,
                    size_t workspace_bytes, int64_t M_max, int64_t *M_out, sgp_stream_t stream);

/* sgp_permute_replay with a row stride: replay_out is [N, stride, 2] with stride >= d+1, the entries past d+1 are
 * {0, 0.0f}.  stride = d+1 rounded up to even gives the slice 16-byte aligned points (sgp_lattice_view.replay_stride). */
int sgp_permute_replay_padded(const int32_t *replay, const uint32_t *perm, const uint32_t *pos, int64_t N, int d,
                              int stride, int32_t *replay_out, sgp_stream_t stream);

/* Test hook: number of fp32 bit patterns a in [lo, lo+count) for which the division-by-constant
 * used inside sgp_slice differs from the IEEE division a / sgp_slice_divisor(d).  Must be 0. */
int sgp_debug_division_mismatches(int d, uint32_t lo, uint32_t count, unsigned long long *mismatches_dev,
                                  sgp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SGP_LATTICE_H */
