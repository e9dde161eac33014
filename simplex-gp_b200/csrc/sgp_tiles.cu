// sgp_tiles.cu -- sorted forms of the point-vertex table for splat and slice (B200, sm_100a).
//
// Three things live here, all built on a radix sort of the N(d+1) point-vertices (CUB, toolkit header library --
// plumbing of the build phase; every kernel on the MVM path is hand-written):
//
// 1. ROW-SORTED SPLAT (production; sgp_build_rowsorted / sgp_splat_rows).  At the metric configuration (N=1M, d=8,
//    16 RHS, M=0.4M) a lattice point is touched by 22 point-vertices on average and by ~9000 at the centre of the data.
//    The scatter splat sends 9M x 64 B vector reductions to L2, where same-address reductions serialise (125 us).  With
//    the point-vertices sorted by lattice row, a thread owns 8 consecutive entries, gathers their RHS rows as plain
//    loads and issues one reduction per run of equal rows: balanced, 6x fewer reductions, 88 us.
//
// 2. LOCALITY ORDER OF THE POINTS (sgp_sort_points, sgp_permute_replay): optional; measured slower than the input
//    order with the plain kernels (adjacent threads then hit the same L2 lines at the same time).
//
// 3. LOCALITY TILES (sgp_tiles_*, sgp_splat_tiles, sgp_slice_tiles): optional.  Points in locality order are cut into
//    tiles of T points; the distinct lattice rows a tile touches form its DICTIONARY (`seg_row`) and its
//    point-vertices are grouped by dictionary entry into SEGMENTS, cut into PIECES of <= 8 entries.  splat: a CTA
//    stages its tile's RHS rows in shared memory and issues one reduction per piece; slice: a CTA stages its tile's
//    dictionary rows once and every point combines its d+1 vertices from shared memory.  L2 row traffic drops 3.3x,
//    but the 9M x 64 B row reads then go through shared memory, whose LSU cost is no lower than the L2 path's at this
//    reuse factor: 101 / 97 us against 88 / 66 us for the production kernels.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok
#define grid_for sgp_grid_for

// ------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgp_tile_pointkeys_kernel(const int32_t *__restrict__ replay, int64_t N, int dp1, uint32_t *__restrict__ keys,
                          uint32_t *__restrict__ vals)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    keys[n] = (uint32_t)replay[n * dp1 * 2];   // lattice index of the remainder-0 vertex
    vals[n] = (uint32_t)n;
}

// composite key of every sorted point-vertex q = p*(d+1)+r: (tile of p) << 32 | lattice index
__global__ void __launch_bounds__(256)
sgp_tile_pvkeys_kernel(const int32_t *__restrict__ replay, const uint32_t *__restrict__ perm, int64_t total, int dp1,
                       int T, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const int64_t p = q / dp1;
    const int r = (int)(q - p * dp1);
    const int64_t n = perm[p];
    const uint32_t idx = (uint32_t)replay[(n * dp1 + r) * 2];
    keys[q] = ((unsigned long long)(p / T) << 32) | idx;
    vals[q] = (uint32_t)q;
}

__global__ void __launch_bounds__(256)
sgp_tile_heads_kernel(const unsigned long long *__restrict__ keys, int64_t total, uint32_t *__restrict__ heads)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    heads[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
sgp_tile_fill_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                     const uint32_t *__restrict__ excl, int64_t total, const int32_t *__restrict__ replay,
                     const uint32_t *__restrict__ perm, int dp1, int T, int64_t n_tiles, int64_t S,
                     uint32_t *__restrict__ seg_ptr, int32_t *__restrict__ seg_row, int2 *__restrict__ seg_ent,
                     uint32_t *__restrict__ tile_seg_ptr, uint32_t *__restrict__ seg_of_q,
                     float *__restrict__ tile_w)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const unsigned long long key = keys[k];
    const bool head = (k == 0) || (key != keys[k - 1]);
    const uint32_t s = excl[k] + (head ? 1u : 0u) - 1u;
    const uint32_t q = vals[k];
    const int64_t p = q / (uint32_t)dp1;
    const int r = (int)(q - p * dp1);
    const int64_t n = perm[p];
    const int32_t wbits = replay[(n * dp1 + r) * 2 + 1];
    const int64_t tile = (int64_t)(key >> 32);
    seg_ent[k] = make_int2((int)(p - tile * T), wbits);   // {point index inside the tile, weight bits}
    seg_of_q[q] = s;
    tile_w[q] = __int_as_float(wbits);
    if (head) {
        seg_ptr[s] = (uint32_t)k;
        seg_row[s] = (int32_t)(uint32_t)key;
        if (k == 0 || (int64_t)(keys[k - 1] >> 32) != tile) tile_seg_ptr[tile] = s;
    }
    if (k == total - 1) {
        seg_ptr[S] = (uint32_t)total;
        tile_seg_ptr[n_tiles] = (uint32_t)S;
    }
}

__global__ void __launch_bounds__(256)
sgp_tile_lidx_kernel(const uint32_t *__restrict__ seg_of_q, const uint32_t *__restrict__ tile_seg_ptr, int64_t total,
                     int dp1, int T, uint16_t *__restrict__ lidx)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const int64_t tile = (q / dp1) / T;
    lidx[q] = (uint16_t)(seg_of_q[q] - tile_seg_ptr[tile]);
}

// per-tile dictionary size -> max (for the slice kernel's shared-memory budget)
__global__ void __launch_bounds__(256)
sgp_tile_maxdict_kernel(const uint32_t *__restrict__ tile_seg_ptr, int64_t n_tiles, uint32_t *__restrict__ max_out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    atomicMax(max_out, tile_seg_ptr[t + 1] - tile_seg_ptr[t]);
}

// workspace layout (bytes, every block 256-aligned):
//   k64a, k64b : total * 8     composite keys (double buffer; also used as 2 x uint32 buffers for the point sort)
//   v32a, v32b : total * 4     payloads
//   heads      : total * 4     head flags -> exclusive scan
//   seg_of_q   : total * 4
//   scan tiles : sgp_scan_tiles(total) * 4
//   total_dev  : 16
//   cub temp   : max of the two sorts
struct TilesWs {
    size_t k64a, k64b, v32a, v32b, heads, seg_of_q, scan_tiles, total_dev, cub, cub_bytes, bytes;
};

static int tiles_ws_layout(int64_t N, int d, TilesWs *w)
{
    const int64_t total = N * (int64_t)(d + 1);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    w->k64a = o; o += al((size_t)total * 8);
    w->k64b = o; o += al((size_t)total * 8);
    w->v32a = o; o += al((size_t)total * 4);
    w->v32b = o; o += al((size_t)total * 4);
    w->heads = o; o += al((size_t)total * 4);
    w->seg_of_q = o; o += al((size_t)total * 4);
    w->scan_tiles = o; o += al((size_t)sgp_scan_tiles(total) * 4);
    w->total_dev = o; o += 256;
    size_t t1 = 0, t2 = 0;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                     (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)N, 0, 32);
    cudaError_t e2 = cub::DeviceRadixSort::SortPairs(nullptr, t2, (const unsigned long long *)nullptr,
                                                     (unsigned long long *)nullptr, (const uint32_t *)nullptr,
                                                     (uint32_t *)nullptr, (int64_t)total, 0, 64);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return sgp_fail(SGP_ECUDA, "cub temp-size query failed");
    w->cub_bytes = t1 > t2 ? t1 : t2;
    w->cub = o; o += al(w->cub_bytes);
    w->bytes = o;
    return SGP_OK;
}

extern "C" size_t sgp_tiles_workspace_bytes(int64_t N, int d)
{
    TilesWs w;
    if (N <= 0 || d < 1 || tiles_ws_layout(N, d, &w) != SGP_OK) return 0;
    return w.bytes;
}

static int bits_for(uint64_t v)
{
    int b = 1;
    while (b < 64 && (v >> b) != 0) ++b;
    return b;
}

extern "C" int sgp_tiles_prepare(const int32_t *replay, int64_t N, int d, int64_t M, int tile_points,
                                 const uint32_t *perm, void *workspace, size_t workspace_bytes, int64_t *S_out,
                                 sgp_stream_t stream)
{
    SGP_RANGE("sgp_tiles_prepare");
    if (!replay || !perm || !workspace || !S_out || N <= 0 || M <= 0 || d < 1 || d > SGP_MAX_DIM)
        return fail(SGP_EINVAL, "sgp_tiles_prepare: bad argument");
    if (tile_points < 1 || (int64_t)tile_points * (d + 1) > 65535)
        return fail(SGP_EINVAL, "tile_points*(d+1) must fit 16 bits (got %d*%d)", tile_points, d + 1);
    TilesWs w;
    int rc = tiles_ws_layout(N, d, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "tiles workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)workspace;
    const int dp1 = d + 1;
    const int64_t total = N * dp1;
    const int64_t n_tiles = (N + tile_points - 1) / tile_points;
    unsigned long long *k64a = (unsigned long long *)(base + w.k64a), *k64b = (unsigned long long *)(base + w.k64b);
    uint32_t *v32a = (uint32_t *)(base + w.v32a), *v32b = (uint32_t *)(base + w.v32b);
    uint32_t *heads = (uint32_t *)(base + w.heads);
    uint32_t *scan_tiles = (uint32_t *)(base + w.scan_tiles);
    unsigned long long *total_dev = (unsigned long long *)(base + w.total_dev);
    void *cub_tmp = base + w.cub;
    size_t cub_bytes = w.cub_bytes;

    // 1. the points arrive sorted (perm from sgp_sort_points)
    // 2. sort point-vertices by (tile, lattice index); stable, so a segment keeps sorted-point order
    sgp_tile_pvkeys_kernel<<<grid_for(total, 256), 256, 0, st>>>(replay, perm, total, dp1, tile_points, k64a, v32a);
    rc = launch_ok("sgp_tile_pvkeys_kernel");
    if (rc) return rc;
    cub_bytes = w.cub_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k64a, k64b, v32a, v32b, (int64_t)total, 0,
                                             32 + bits_for((uint64_t)n_tiles), st));
    // 3. segment heads -> exclusive scan -> S
    sgp_tile_heads_kernel<<<grid_for(total, 256), 256, 0, st>>>(k64b, total, heads);
    rc = launch_ok("sgp_tile_heads_kernel");
    if (rc) return rc;
    rc = sgp_exclusive_scan_u32(heads, total, scan_tiles, total_dev, st);
    if (rc) return rc;
    unsigned long long s_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&s_host, total_dev, sizeof(s_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *S_out = (int64_t)s_host;
    return SGP_OK;
}

extern "C" int sgp_tiles_finalize(const int32_t *replay, const uint32_t *perm, int64_t N, int d, int tile_points,
                                  int64_t S, void *workspace, size_t workspace_bytes, uint32_t *seg_ptr,
                                  int32_t *seg_row, int32_t *seg_ent, uint32_t *tile_seg_ptr, uint16_t *lidx,
                                  float *tile_w, int32_t *max_dict_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_tiles_finalize");
    if (!replay || !perm || !workspace || !seg_ptr || !seg_row || !seg_ent || !tile_seg_ptr || !lidx || !tile_w ||
        !max_dict_out || N <= 0 || S <= 0)
        return fail(SGP_EINVAL, "sgp_tiles_finalize: bad argument");
    TilesWs w;
    int rc = tiles_ws_layout(N, d, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "tiles workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)workspace;
    const int dp1 = d + 1;
    const int64_t total = N * dp1;
    const int64_t n_tiles = (N + tile_points - 1) / tile_points;
    const unsigned long long *keys = (const unsigned long long *)(base + w.k64b);
    const uint32_t *vals = (const uint32_t *)(base + w.v32b);
    const uint32_t *excl = (const uint32_t *)(base + w.heads);
    uint32_t *seg_of_q = (uint32_t *)(base + w.seg_of_q);
    uint32_t *max_dev = (uint32_t *)(base + w.total_dev);
    sgp_tile_fill_kernel<<<grid_for(total, 256), 256, 0, st>>>(keys, vals, excl, total, replay, perm, dp1, tile_points,
                                                                n_tiles, S, seg_ptr, seg_row, (int2 *)seg_ent,
                                                                tile_seg_ptr, seg_of_q, tile_w);
    rc = launch_ok("sgp_tile_fill_kernel");
    if (rc) return rc;
    sgp_tile_lidx_kernel<<<grid_for(total, 256), 256, 0, st>>>(seg_of_q, tile_seg_ptr, total, dp1, tile_points, lidx);
    rc = launch_ok("sgp_tile_lidx_kernel");
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(max_dev, 0, sizeof(uint32_t), st));
    sgp_tile_maxdict_kernel<<<grid_for(n_tiles, 256), 256, 0, st>>>(tile_seg_ptr, n_tiles, max_dev);
    rc = launch_ok("sgp_tile_maxdict_kernel");
    if (rc) return rc;
    uint32_t mx = 0;
    CUDA_TRY(cudaMemcpyAsync(&mx, max_dev, sizeof(mx), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *max_dict_out = (int32_t)mx;
    return SGP_OK;
}

// ------------------------------------------------------------------------------------
// locality order of the points: lexicographic by the remainder-0 lattice point (greedy), so that points of one
// lattice cell -- and, mostly, of neighbouring cells -- are adjacent.  LSD radix sort, four 16-bit coordinates per pass.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgp_pointsort_keys_kernel(const int16_t *__restrict__ greedy, const uint32_t *__restrict__ perm_in, int64_t N, int dp1,
                          int c_hi, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const uint32_t n = perm_in ? perm_in[p] : (uint32_t)p;
    const int16_t *g = greedy + (int64_t)n * dp1;
    unsigned long long k = 0;
    // coordinates c_hi-3 .. c_hi, most significant first (c_hi-3 is the more significant axis)
    for (int c = c_hi - 3; c <= c_hi; ++c) {
        const uint32_t v = (c >= 0) ? (uint32_t)(uint16_t)(g[c] + 32768) : 0u;
        k = (k << 16) | v;
    }
    keys[p] = k;
    vals[p] = n;
}

extern "C" size_t sgp_sort_points_workspace_bytes(int64_t N)
{
    if (N <= 0) return 0;
    size_t t = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, t, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)N, 0, 64) != cudaSuccess)
        return 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al((size_t)N * 8) * 2 + al((size_t)N * 4) * 2 + al(t);
}

extern "C" int sgp_sort_points(const int16_t *greedy, int64_t N, int d, uint32_t *perm, void *workspace,
                               size_t workspace_bytes, sgp_stream_t stream)
{
    SGP_RANGE("sgp_sort_points");
    if (!greedy || !perm || !workspace || N <= 0 || d < 1 || d > SGP_MAX_DIM) return fail(SGP_EINVAL, "sgp_sort_points: bad argument");
    if (workspace_bytes < sgp_sort_points_workspace_bytes(N)) return fail(SGP_EINVAL, "sgp_sort_points: workspace too small");
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char *base = (char *)workspace;
    unsigned long long *ka = (unsigned long long *)base, *kb = (unsigned long long *)(base + al((size_t)N * 8));
    uint32_t *va = (uint32_t *)(base + 2 * al((size_t)N * 8)), *vb = va + al((size_t)N * 4) / 4;
    void *tmp = (char *)(vb) + al((size_t)N * 4);
    size_t tmp_bytes = workspace_bytes - (size_t)((char *)tmp - base);
    cudaStream_t st = (cudaStream_t)stream;
    const int dp1 = d + 1;
    const uint32_t *cur = nullptr;   // identity
    // least significant group of coordinates first; sorts are stable, so the final order is lexicographic in 0..d
    for (int c_hi = d; c_hi >= 0; c_hi -= 4) {
        sgp_pointsort_keys_kernel<<<grid_for(N, 256), 256, 0, st>>>(greedy, cur, N, dp1, c_hi, ka, va);
        int rc = launch_ok("sgp_pointsort_keys_kernel");
        if (rc) return rc;
        uint32_t *dst = (c_hi - 4 < 0) ? perm : vb;
        size_t tb = tmp_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, ka, kb, va, dst, (int64_t)N, 0, 64, st));
        if (dst != perm) {   // next pass reads its input permutation from vb
            cur = vb;
        }
    }
    return SGP_OK;
}

// out[p, r] = {pos ? pos[replay[perm[p], r].index] : replay[perm[p], r].index, weight bits}
__global__ void __launch_bounds__(256)
sgp_permute_replay_kernel(const int2 *__restrict__ replay, const uint32_t *__restrict__ perm,
                          const uint32_t *__restrict__ pos, int64_t N, int dp1, int transposed, int2 *__restrict__ out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output index
    if (q >= N * dp1) return;
    int64_t p;
    int r;
    if (transposed) {
        r = (int)(q / N);
        p = q - (int64_t)r * N;
    } else {
        p = q / dp1;
        r = (int)(q - p * dp1);
    }
    int2 e = replay[(perm ? (int64_t)perm[p] : p) * dp1 + r];
    if (pos) e.x = (int)pos[e.x];
    out[q] = e;
}

extern "C" int sgp_permute_replay(const int32_t *replay, const uint32_t *perm, const uint32_t *pos, int64_t N, int d,
                                  int transposed, int32_t *replay_out, sgp_stream_t stream)
{
    if (N == 0) return SGP_OK;
    if (!replay || !replay_out || N < 0 || d < 1) return fail(SGP_EINVAL, "sgp_permute_replay: bad argument");
    const int64_t total = N * (int64_t)(d + 1);
    sgp_permute_replay_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const int2 *)replay, perm, pos, N,
                                                                                       d + 1, transposed, (int2 *)replay_out);
    return launch_ok("sgp_permute_replay_kernel");
}

__global__ void __launch_bounds__(256)
sgp_permute_replay_padded_kernel(const int2 *__restrict__ replay, const uint32_t *__restrict__ perm,
                                 const uint32_t *__restrict__ pos, int64_t N, int dp1, int stride, int2 *__restrict__ out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output index
    if (q >= N * stride) return;
    const int64_t p = q / stride;
    const int r = (int)(q - p * stride);
    int2 e = make_int2(0, 0);
    if (r < dp1) {
        e = replay[(perm ? (int64_t)perm[p] : p) * dp1 + r];
        if (pos) e.x = (int)pos[e.x];
    }
    out[q] = e;
}

extern "C" int sgp_permute_replay_padded(const int32_t *replay, const uint32_t *perm, const uint32_t *pos, int64_t N, int d,
                                         int stride, int32_t *replay_out, sgp_stream_t stream)
{
    if (N == 0) return SGP_OK;
    if (!replay || !replay_out || N < 0 || d < 1 || stride < d + 1) return fail(SGP_EINVAL, "sgp_permute_replay_padded: bad argument");
    const int64_t total = N * (int64_t)stride;
    sgp_permute_replay_padded_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const int2 *)replay, perm, pos, N,
                                                                                              d + 1, stride, (int2 *)replay_out);
    return launch_ok("sgp_permute_replay_padded_kernel");
}

// dst[n, 0..Lv) = src[n, 0..L) followed by zeros (Lv % 4 == 0, dst 16-byte aligned with ldd % 4 == 0): the zero-padded
// copy of a ragged right-hand-side block that lets the splat gather 16-byte vectors (SGP_MVM_SRC_PADDED).  One thread
// per 16-byte piece of dst: coalesced scalar reads of the contiguous source rows, one vector store.
__global__ void __launch_bounds__(256)
sgp_pad_columns_kernel(const float *__restrict__ src, int64_t lds, int L, float *__restrict__ dst, int64_t ldd, int Lv, int64_t N)
{
    const int pieces = Lv >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * pieces) return;
    const int64_t n = t / pieces;
    const int c = (int)(t - n * pieces) << 2;
    const float *s = src + n * lds + c;
    float4 v;
    v.x = c + 0 < L ? __ldcs(s + 0) : 0.0f;
    v.y = c + 1 < L ? __ldcs(s + 1) : 0.0f;
    v.z = c + 2 < L ? __ldcs(s + 2) : 0.0f;
    v.w = c + 3 < L ? __ldcs(s + 3) : 0.0f;
    *(float4 *)(dst + n * ldd + c) = v;
}

extern "C" int sgp_pad_columns(const float *src, int64_t lds, int L, float *dst, int64_t ldd, int Lv, int64_t N,
                               sgp_stream_t stream)
{
    if (N == 0) return SGP_OK;
    if (!src || !dst || N < 0 || L < 1 || Lv < L || Lv % 4 != 0 || lds < L || ldd < Lv || ldd % 4 != 0 ||
        ((uintptr_t)dst & 15) != 0)
        return fail(SGP_EINVAL, "sgp_pad_columns: bad argument");
    const int64_t work = N * (int64_t)(Lv / 4);
    sgp_pad_columns_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, L, dst, ldd, Lv, N);
    return launch_ok("sgp_pad_columns_kernel");
}

// ------------------------------------------------------------------------------------
// Row-sorted splat ("segmented gather").  The point-vertices are sorted once per lattice by the lattice row they
// touch (stable radix sort: within a row they stay in point-vertex order, the reference's accumulation order).  A
// thread then owns ROWSEG consecutive entries and one channel chunk: it issues the ROWSEG RHS-row loads together,
// accumulates runs of equal row in registers and issues one vector reduction per run.  Work is perfectly balanced
// whatever the row lengths (1 to ~10^4 touches per lattice point at the metric configuration); reductions drop from
// one per point-vertex to one per (row, thread) run; and the gather side moves through L2 as plain loads, which the
// scatter form's same-address reductions do not.
// ------------------------------------------------------------------------------------
#ifndef SGP_LDCS
#define SGP_LDCS 1
#endif
#define ROWSEG SGP_ENTRY_GROUP   /* padding granularity of the row-sorted arrays (sgp_entry_index, sgp_common.cuh) */

// q < total: a point-vertex, keyed by its lattice row.  total <= q < total + fill: one weightless filler per lattice
// row 0..fill-1 (value 0xFFFFFFFF), for lattices whose points do not touch every row (a rank's share of the points
// on the full key set): the row-start encoding below needs every row to own at least one entry.
#define SGP_ROW_FILLER 0xFFFFFFFFu

__global__ void __launch_bounds__(256)
sgp_rowsort_keys_kernel(const int2 *__restrict__ replay, int64_t total, int64_t fill, uint32_t *__restrict__ keys,
                        uint32_t *__restrict__ vals)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total + fill) return;
    keys[q] = q < total ? (uint32_t)replay[q].x : (uint32_t)(q - total);
    vals[q] = q < total ? (uint32_t)q : SGP_ROW_FILLER;
}

// ent[k] = {point | SGP_ROW_START, weight}: the flag marks the first entry of a lattice row.  Every lattice row has at
// least one entry and the rows are consecutive integers, so a thread that knows the row of its first entry
// (seg_row, one int per 4 entries) finds the others by counting flags: 8.5 bytes per entry instead of 12.
#define SGP_ROW_START 0x80000000u
#define SGP_ROW_GRAIN 4

__global__ void __launch_bounds__(256)
sgp_rowsort_fill_kernel(const int2 *__restrict__ replay, const uint32_t *__restrict__ sorted_row,
                        const uint32_t *__restrict__ sorted_pv, int64_t total, int64_t padded, int dp1,
                        int2 *__restrict__ ent, int32_t *__restrict__ ent_row, int32_t *__restrict__ seg_row)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= padded) return;
    uint32_t row;
    if (k < total) {
        const uint32_t q = sorted_pv[k];
        row = sorted_row[k];
        const uint32_t start = (k > 0 && sorted_row[k - 1] != row) ? SGP_ROW_START : 0u;
        ent[sgp_entry_index(k)] = q == SGP_ROW_FILLER ? make_int2((int)start, 0)
                                                      : make_int2((int)((q / (uint32_t)dp1) | start), replay[q].y);
    } else {   // padding: weight 0 on the last row
        row = sorted_row[total - 1];
        ent[sgp_entry_index(k)] = make_int2(0, 0);
    }
    if (ent_row) ent_row[k] = (int32_t)row;
    if (k % SGP_ROW_GRAIN == 0) seg_row[k / SGP_ROW_GRAIN] = (int32_t)row;
}

extern "C" size_t sgp_rowsort_workspace_bytes(int64_t N, int d, int64_t fill_rows)
{
    const int64_t total = N * (int64_t)(d + 1) + (fill_rows > 0 ? fill_rows : 0);
    if (total <= 0) return 0;
    size_t t = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                        (uint32_t *)nullptr, (int64_t)total, 0, 32) != cudaSuccess)
        return 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al((size_t)total * 4) * 4 + al(t);
}

extern "C" int64_t sgp_rowsort_padded(int64_t N, int d, int64_t fill_rows)
{
    const int64_t total = N * (int64_t)(d + 1) + (fill_rows > 0 ? fill_rows : 0);
    return (total + ROWSEG - 1) / ROWSEG * ROWSEG;
}

extern "C" int sgp_build_rowsorted(const int32_t *replay, int64_t N, int d, int64_t M, int64_t fill_rows, int32_t *ent,
                                   int32_t *ent_row, int32_t *seg_row, void *workspace, size_t workspace_bytes,
                                   sgp_stream_t stream)
{
    SGP_RANGE("sgp_build_rowsorted");
    if (!replay || !ent || !seg_row || !workspace || N <= 0 || M <= 0 || d < 1) return fail(SGP_EINVAL, "sgp_build_rowsorted: bad argument");
    if (fill_rows != 0 && fill_rows != M) return fail(SGP_EINVAL, "sgp_build_rowsorted: fill_rows must be 0 or M");
    if (N >= (1ll << 31)) return fail(SGP_EOVERFLOW, "sgp_build_rowsorted: N does not fit 31 bits");
    if (workspace_bytes < sgp_rowsort_workspace_bytes(N, d, fill_rows)) return fail(SGP_EINVAL, "sgp_build_rowsorted: workspace too small");
    const int64_t total = N * (int64_t)(d + 1) + fill_rows;
    if (total >= (1ll << 32) - 1) return fail(SGP_EOVERFLOW, "sgp_build_rowsorted: too many entries");
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char *base = (char *)workspace;
    const size_t blk = al((size_t)total * 4);
    uint32_t *ka = (uint32_t *)base, *kb = (uint32_t *)(base + blk), *va = (uint32_t *)(base + 2 * blk),
             *vb = (uint32_t *)(base + 3 * blk);
    void *tmp = base + 4 * blk;
    size_t tmp_bytes = workspace_bytes - 4 * blk;
    cudaStream_t st = (cudaStream_t)stream;
    sgp_rowsort_keys_kernel<<<grid_for(total, 256), 256, 0, st>>>((const int2 *)replay, total - fill_rows, fill_rows, ka, va);
    int rc = launch_ok("sgp_rowsort_keys_kernel");
    if (rc) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ka, kb, va, vb, (int64_t)total, 0, bits_for((uint64_t)M), st));
    const int64_t padded = sgp_rowsort_padded(N, d, fill_rows);
    sgp_rowsort_fill_kernel<<<grid_for(padded, 256), 256, 0, st>>>((const int2 *)replay, kb, vb, total, padded, d + 1,
                                                                  (int2 *)ent, ent_row, seg_row);
    return launch_ok("sgp_rowsort_fill_kernel");
}

// thread = (segment of SEG entries, channel chunk)
// RAGGED: the lattice rows are L (a multiple of VEC) channels wide but src has only L_src < L columns and arbitrary
// row alignment (e.g. the 11-column CG block of a training step, padded to 12 on the lattice): src is read one
// channel at a time, the missing channels as zeros; everything downstream moves 16-byte vectors.
template <int VEC, int SEG, bool RAGGED>
__global__ void __launch_bounds__(256)
sgp_splat_rows_kernel(const int2 *__restrict__ ent, const int32_t *__restrict__ seg_row, int64_t n_seg,
                      const float *__restrict__ src, int64_t lds, int L, int L_src, int chunks, int64_t prefetch_bytes,
                      bool aggregate, float *__restrict__ values, int dbg)
{
    // dbg (experiments only, SGP_SPLAT_DBG): bit 0 = issue no reductions; bits 8-15 = k > 0: point indices folded to k bits
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // Optional (off by default, measured without effect on B200): the first blocks ask L2 for the whole of src with
    // sequential 128-byte prefetches before the random row gathers of the later blocks need it.
    if (prefetch_bytes > 0) {
        const int64_t line = tid * 128;
        if (line < prefetch_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)src + line));
    } else if (prefetch_bytes < 0) {
        // bulk form (TMA engine, no LSU work): the first CTAs each ask L2 for one 32 KB piece of src
        const int64_t piece = 32768, off = (int64_t)blockIdx.x * piece;
        if (threadIdx.x == 0 && off < -prefetch_bytes) {
            const int64_t n = min(piece, -prefetch_bytes - off) & ~(int64_t)15;
            if (n > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)src + off), "r"((uint32_t)n) : "memory");
        }
    }
    pdl_launch_dependents();
    const int64_t seg = tid / chunks;
    const unsigned mask = __ballot_sync(0xffffffffu, seg < n_seg);   // the lanes that take part in the shuffles below
    if (seg >= n_seg) return;
    const int c0 = (int)(tid - seg * chunks) * VEC;
    int2 e[SEG];
    Vec<VEC> v[SEG];
    static_assert(SEG == 8, "the entry layout interleaves segments of 8 entries");
    // piece i of segment seg: the segments of a group of 8 are interleaved piece by piece (sgp_entry_index)
    const int4 *ep = (const int4 *)ent + ((seg >> 3) << 5) + (seg & 7);
    int row = __ldg(seg_row + seg * (SEG / SGP_ROW_GRAIN));   // lattice row of the first entry
#pragma unroll
    for (int i = 0; i < SEG / 2; ++i) {
        const int4 t = SGP_LDCS ? __ldcs(ep + 8 * i) : __ldg(ep + 8 * i);   // read once: streaming, keeps src and the lattice in L2
        e[2 * i] = make_int2(t.x, t.y);
        e[2 * i + 1] = make_int2(t.z, t.w);
    }
    // no weight has all bits set: the branch keeps every entry load ahead of every row load (in-order issue)
    int all = e[0].y;
#pragma unroll
    for (int i = 1; i < SEG; ++i) all &= e[i].y;
    if (all == -1 || row < 0) return;
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        if (RAGGED) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                v[i].v[k] = (c0 + k < L_src) ? ldg_ordered_f1(src + (int64_t)(e[i].x & 0x7fffffff) * lds + c0 + k) : 0.0f;
        } else {
            v[i].load_ordered(src + (int64_t)(e[i].x & 0x7fffffff & (dbg >> 8 ? (1 << (dbg >> 8)) - 1 : 0x7fffffff)) * lds + c0);
        }
    }
    Vec<VEC> acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
    // A thread whose SEG entries all belong to one lattice row ("single") does not reduce into memory by itself: the
    // most-touched rows span hundreds of consecutive segments (1250 at the centre of the metric lattice), and that
    // many reductions to one address serialise in L2.  Singles of the same row and channel chunk that sit in the same
    // warp are summed with shuffles first and the first of them issues one reduction.
    bool single = aggregate || (dbg & 4);
#pragma unroll
    for (int i = 1; i < SEG; ++i) single = single && !(e[i].x < 0);
    // Warp-uniform case (dbg bit 2, the default for 3+ channel chunks): all segments of the warp lie inside ONE lattice
    // row -- the long rows at the centre of the data, where most reductions go.  Their partial sums are added across the
    // warp's segments with a butterfly and the first segment issues one reduction per channel chunk: 8x fewer L2
    // atomics for those warps, and nothing but one vote and one shuffle for all the others.
    bool uniform = false;
    if ((dbg & 4) && !aggregate) {
        const int row_first = __shfl_sync(mask, row, 0);
        uniform = mask == 0xffffffffu && __all_sync(mask, single && row == row_first) && (32 % chunks == 0);
        single = uniform;   // non-uniform warps: every thread reduces its own runs, as before
    }
    pdl_wait();   // the lattice values are zeroed (and last read) by the stream's previous work
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        const float w = __int_as_float(e[i].y);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = __fmaf_rn(w, v[i].v[k], acc.v[k]);
        if ((i == SEG - 1 && !single) || (i < SEG - 1 && e[i + 1].x < 0)) {      // the next entry starts the next lattice row
            if (!(dbg & 1) || acc.v[0] == 1.2345e-30f) acc.red(values + (int64_t)row * L + c0);
            ++row;
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
        }
    }
    if (uniform) {
        for (int step = chunks; step < 32; step <<= 1) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] += __shfl_xor_sync(0xffffffffu, acc.v[k], step);
        }
        if ((threadIdx.x & 31) < chunks) acc.red(values + (int64_t)row * L + c0);
        return;
    }
    if (!aggregate) return;
    // segmented sum over the lanes lane, lane + chunks, lane + 2 chunks, ... (same channel chunk, consecutive segments)
    const int lane = threadIdx.x & 31;
    const int key = single ? row : -1 - lane;      // non-singles never match anybody
    const int prev_key = __shfl_up_sync(mask, key, chunks);
    const bool prev_there = lane >= chunks && ((mask >> (lane - chunks)) & 1u);
    if (__any_sync(mask, single && prev_there && prev_key == key)) {
        for (int step = chunks; step < 32; step <<= 1) {
            const int other_key = __shfl_down_sync(mask, key, step);
            const bool take = single && lane + step < 32 && ((mask >> (lane + step)) & 1u) && other_key == key;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float y = __shfl_down_sync(mask, acc.v[k], step);
                if (take) acc.v[k] += y;
            }
        }
    }
    if (single && !(prev_there && prev_key == key)) acc.red(values + (int64_t)row * L + c0);
}

static int splat_rows_impl(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                           const float *src, int64_t lds, int L_src, float *values, int L, bool prezeroed,
                           sgp_stream_t stream);

extern "C" int sgp_splat_rows(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                              const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream)
{
    return splat_rows_impl(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, false, stream);
}

extern "C" int sgp_mvm_stage_splat_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N,
                                             int64_t M, const float *src, int64_t lds, int L_src, float *values, int L,
                                             sgp_stream_t stream)
{
    return splat_rows_impl(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, true, stream);
}

// values must hold zeros on entry: the caller zeroed it off the critical path (sgp_mvm_rows_groups_ex)
int sgp_splat_rows_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                             const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream)
{
    return splat_rows_impl(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, true, stream);
}

static int splat_rows_impl(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                           const float *src, int64_t lds, int L_src, float *values, int L, bool prezeroed,
                           sgp_stream_t stream)
{
    SGP_RANGE("sgp_splat_rows");
    if (N == 0 || M == 0) return SGP_OK;
    if (!ent || !seg_row || !src || !values || N < 0 || M < 0 || n_entries < 0 || n_entries % ROWSEG != 0 || L_src < 1 ||
        lds < L_src || L < L_src)
        return fail(SGP_EINVAL, "sgp_splat_rows: bad argument");
    // production form for dense lattices: index stream through warp-private TMA rings (sgp_ring.cu); SGP_RING=0 or
    // SGP_RING_SPLAT=0 selects the one-shot kernel below
    // ... where it wins: long rows (a dense lattice: the metric shape has 22 entries per lattice row) and enough tiles to
    // keep every persistent warp busy, and 2 to 16 float4 chunks per row (3, 5, 6, 7, 9+ chunks run on power-of-two lane
    // slots with idle lanes, which keeps the warp-uniform aggregation).  Measured (profiles/exp_splat_L.py, metric
    // lattice, one-shot -> ring): L = 8: 59.5 -> 57.8 us, 12: 83.4 -> 75.3, 16: 86.5 -> 73.6, 20: 126.1 -> 111.0,
    // 24: 142.0 -> 121.4, 28: 176.0 -> 150.4, 32: 160.8 -> 120.0, 40: 241.5 -> 219.6; but L = 1: 42.2 -> 49.4,
    // L = 4: 49.6 -> 53.7; config C (1.5 entries per row) 434 -> 510 us and config B (16.6 k points) 8 -> 19 us: those
    // keep the one-shot kernel.  SGP_RING_FORCE=1 overrides.
    const bool ring_pays = (n_entries >= 8 * M && n_entries >= (1 << 21) && L % 4 == 0 && L >= 8 && L <= 64) ||
                           (getenv("SGP_RING_FORCE") && atoi(getenv("SGP_RING_FORCE")));
    if (sgp_ring_splat_enabled() && ring_pays && n_entries >= 64 && sgp_splat_ring_supported(values, L))
        return prezeroed ? sgp_splat_rows_ring_prezeroed(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, stream)
                         : sgp_splat_rows_ring(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int seg_env = 8;
    const int64_t n_seg = n_entries / seg_env;
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    // the lattice side decides the vector width; src is read channel by channel when it does not match it
    int vec = 1;
    if (L % 4 == 0 && al(values, 16)) vec = 4;
    else if (L % 2 == 0 && al(values, 8)) vec = 2;
    const bool ragged = vec > 1 && !(L_src == L && lds % vec == 0 && al(src, 4 * vec));
    const int chunks = L / vec;
    const char *dbg_e = getenv("SGP_SPLAT_DBG");   // experiments only: bit 0 no reductions, bit 1 no memset, bits 8+ fold
    int dbg = dbg_e ? atoi(dbg_e) : 0;
    if (!(dbg & 2) && !prezeroed) CUDA_TRY(cudaMemsetAsync(values, 0, sizeof(float) * (size_t)M * (size_t)L, st));
    const int64_t work = n_seg * chunks;
    int pref_env = 0;   // tuning hook: SGP_SPLAT_PREFETCH=1|2 asks L2 for src up front (read on every call)
    {
        const char *e = getenv("SGP_SPLAT_PREFETCH");
        pref_env = e ? atoi(e) : 0;
    }
    // SGP_SPLAT_PREFETCH: 1 = one prefetch.global.L2 per 128-byte line, 2 = cp.async.bulk.prefetch.L2 in 32 KB pieces
    // (src must then be 16-byte aligned)
    int64_t prefetch_bytes = pref_env ? (int64_t)N * lds * (int64_t)sizeof(float) : 0;
    if (pref_env == 2) prefetch_bytes = al(src, 16) ? -prefetch_bytes : 0;
    // warp-level aggregation of whole-segment runs: lanes l, l + chunks, l + 2 chunks, ... hold the same channel chunk of
    // consecutive segments
    static int agg_env = -1;    // tuning hook: SGP_SPLAT_AGG=0 disables
    if (agg_env < 0) {
        const char *e = getenv("SGP_SPLAT_AGG");
        agg_env = e ? atoi(e) : 1;
    }
    // Measured at the metric shape: 1 column 55 -> 43 us, 2 columns 54 -> 46, 4 columns 57 -> 52 (one or two chunks: up to
    // 32 / 16 segments of a row meet in a warp); neutral at 12 columns, 2 us slower at 16 (8 segments per warp do not pay
    // for the shuffles), so it is used for one or two chunks only.
    const bool aggregate = agg_env && chunks <= 2;
    {
        const char *e = getenv("SGP_SPLAT_UAGG");   // warp-uniform aggregation for wide rows (default on; power-of-two chunk counts)
        const int uagg = e ? atoi(e) : 1;
        if (uagg && !aggregate && (chunks & (chunks - 1)) == 0 && !(dbg & 1)) dbg |= 4;
    }
    cudaError_t launch_err = cudaSuccess;
#define SGP_ROWS_LAUNCH(VV, SS)                                                                                        \
    launch_err = ragged ? sgp_launch_pdl(sgp_splat_rows_kernel<VV, SS, true>, dim3(grid_for(work, 256)), dim3(256), 0,  \
                                         st, (const int2 *)ent, seg_row, n_seg, src, lds, L, L_src, chunks,            \
                                         prefetch_bytes, aggregate, values, dbg)                                      \
                        : sgp_launch_pdl(sgp_splat_rows_kernel<VV, SS, false>, dim3(grid_for(work, 256)), dim3(256), 0, \
                                         st, (const int2 *)ent, seg_row, n_seg, src, lds, L, L_src, chunks,            \
                                         prefetch_bytes, aggregate, values, dbg)
#define SGP_ROWS_SEG(VV) SGP_ROWS_LAUNCH(VV, 8)
    if (vec == 4) SGP_ROWS_SEG(4);
    else if (vec == 2) SGP_ROWS_SEG(2);
    else SGP_ROWS_SEG(1);
#undef SGP_ROWS_SEG
#undef SGP_ROWS_LAUNCH
    if (launch_err != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_splat_rows_kernel failed: %s", cudaGetErrorString(launch_err));
    return launch_ok("sgp_splat_rows_kernel");
}

// ------------------------------------------------------------------------------------
// MVM kernels on tiles.  One CTA per tile and per block of CB channels (CB*4 = one 64-byte row piece for L >= 16).
// Every global read of a tile happens in a first phase of independent, coalesced (or row-gather) loads into shared
// memory; the second phase works out of shared memory only, so L2 sees one row transfer per dictionary entry
// (slice) or per piece (splat) instead of one per point-vertex.
// ------------------------------------------------------------------------------------
#define TILE_THREADS 256
#define TILE_PIECE 8   /* entries per splat piece: bounds the work of one thread and the length of one dependent chain */

// splat.  Shared memory: V[T][CB] rows of the tile's points | E[T*dp1] {local point, weight} grouped by piece |
// PP[pieces+1] piece bounds (relative to the tile's first entry) | PR[pieces] lattice row of each piece.
template <int VEC>
__global__ void __launch_bounds__(TILE_THREADS)
sgp_splat_tiles_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ tile_piece_ptr,
                       const uint32_t *__restrict__ piece_ptr, const int32_t *__restrict__ piece_row,
                       const int2 *__restrict__ seg_ent, const float *__restrict__ src, int64_t lds, int64_t N, int T,
                       int dp1, int L, int CB, float *__restrict__ values)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // T is a multiple of 8, so every array below starts 16-byte aligned
    float *V = (float *)smem_raw;
    int2 *E = (int2 *)(V + (size_t)T * CB);
    uint32_t *PP = (uint32_t *)(E + (size_t)T * dp1);
    int32_t *PR = (int32_t *)(PP + (size_t)T * dp1 + 8);
    uint32_t *PM = (uint32_t *)(PR + (size_t)T * dp1);   // the tile's slice of perm
    const int64_t tile = blockIdx.x;
    const int cb0 = blockIdx.y * CB;
    const int cb = min(CB, L - cb0);
    const int chunks = cb / VEC;
    const int64_t p0 = tile * T;
    const int np = (int)min((int64_t)T, N - p0);
    const int64_t e0 = p0 * dp1;
    const uint32_t j0 = tile_piece_ptr[tile];
    const int npieces = (int)(tile_piece_ptr[tile + 1] - j0);

    // phase 0a: everything that is a plain range copy, asynchronously (perm slice, entries, piece bounds and rows)
    cta_copy_async(PM, perm + p0, np * 4, threadIdx.x, TILE_THREADS);
    cta_copy_async(E, seg_ent + e0, np * dp1 * 8, threadIdx.x, TILE_THREADS);
    cta_copy_async(PP, piece_ptr + j0, (npieces + 1) * 4, threadIdx.x, TILE_THREADS);
    cta_copy_async(PR, piece_row + j0, npieces * 4, threadIdx.x, TILE_THREADS);
    cp_async_wait_all();
    __syncthreads();
    // phase 0b: the RHS rows of the tile's points, gathered through perm
    for (int w = threadIdx.x; w < np * chunks; w += TILE_THREADS) {
        const int lp = w / chunks, c = (w - lp * chunks) * VEC;
        cp_async_vec<VEC>(V + lp * CB + c, src + (int64_t)PM[lp] * lds + cb0 + c);
    }
    cp_async_wait_all();
    __syncthreads();

    for (int w = threadIdx.x; w < npieces * chunks; w += TILE_THREADS) {
        const int j = w / chunks, c = (w - j * chunks) * VEC;
        const uint32_t a = PP[j] - (uint32_t)e0, b = PP[j + 1] - (uint32_t)e0;
        Vec<VEC> acc;
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
#pragma unroll 4
        for (uint32_t e = a; e < b; ++e) {
            const int2 ent = E[e];
            const float wgt = __int_as_float(ent.y);
            Vec<VEC> sv;
            sv.load_plain(V + ent.x * CB + c);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] = __fmaf_rn(wgt, sv.v[k], acc.v[k]);
        }
        acc.red(values + (int64_t)PR[j] * L + cb0 + c);
    }
}

// slice.  Shared memory: D[cap][CB] dictionary rows | W[T*dp1] weights | LI[T*dp1] dictionary index per point-vertex.
// Tiles whose dictionary exceeds cap rows (sparse regions: no reuse to exploit) read the lattice directly.
template <int VEC, bool FAST>
__global__ void __launch_bounds__(TILE_THREADS)
sgp_slice_tiles_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ tile_seg_ptr,
                       const int32_t *__restrict__ seg_row, const uint16_t *__restrict__ lidx,
                       const float *__restrict__ tile_w, const float *__restrict__ values, int64_t N, int T, int dp1,
                       int L, int CB, int cap, float divisor, float rdivisor, float *__restrict__ out, int64_t ldo)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *D = (float *)smem_raw;
    float *W = D + (((size_t)cap * CB + 3) & ~(size_t)3);
    int32_t *IDX = (int32_t *)(W + (size_t)T * dp1);          // the tile's dictionary (lattice rows)
    uint32_t *PM = (uint32_t *)(IDX + ((cap + 3) & ~3));      // the tile's slice of perm
    uint16_t *LI = (uint16_t *)(PM + T);
    const int64_t tile = blockIdx.x;
    const int cb0 = blockIdx.y * CB;
    const int cb = min(CB, L - cb0);
    const int chunks = cb / VEC;
    const int64_t p0 = tile * T;
    const int np = (int)min((int64_t)T, N - p0);
    const int64_t q0 = p0 * dp1;
    const uint32_t s0 = tile_seg_ptr[tile];
    const int nloc = (int)(tile_seg_ptr[tile + 1] - s0);
    const bool staged = nloc <= cap;

    // phase 0a: plain range copies, asynchronously (dictionary, weights, dictionary indices, perm slice)
    if (staged) cta_copy_async(IDX, seg_row + s0, nloc * 4, threadIdx.x, TILE_THREADS);
    cta_copy_async(W, tile_w + q0, np * dp1 * 4, threadIdx.x, TILE_THREADS);
    cta_copy_async(LI, lidx + q0, (np * dp1 * 2 + 3) & ~3, threadIdx.x, TILE_THREADS);   // lidx is padded by the builder
    cta_copy_async(PM, perm + p0, np * 4, threadIdx.x, TILE_THREADS);
    cp_async_wait_all();
    __syncthreads();
    // phase 0b: the dictionary rows, all in flight at once
    if (staged) {
        for (int w = threadIdx.x; w < nloc * chunks; w += TILE_THREADS) {
            const int lr = w / chunks, c = (w - lr * chunks) * VEC;
            cp_async_vec<VEC>(D + lr * CB + c, values + (int64_t)IDX[lr] * L + cb0 + c);
        }
        cp_async_wait_all();
        __syncthreads();
    }

    for (int w = threadIdx.x; w < np * chunks; w += TILE_THREADS) {
        const int lp = w / chunks, c = (w - lp * chunks) * VEC;
        const float *wp = W + lp * dp1;
        const uint16_t *lp_i = LI + lp * dp1;
        Vec<VEC> acc;
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
        if (staged) {
#pragma unroll 3
            for (int r = 0; r < dp1; ++r) {
                const float wgt = wp[r];
                Vec<VEC> sv;
                sv.load_plain(D + (int)lp_i[r] * CB + c);
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc.v[k] = FAST ? __fmaf_rn(wgt, sv.v[k], acc.v[k])
                                    : __fadd_rn(acc.v[k], exact_div(__fmul_rn(wgt, sv.v[k]), divisor, rdivisor));
            }
        } else {
            for (int r = 0; r < dp1; ++r) {
                const float wgt = wp[r];
                Vec<VEC> v;
                v.load(values + (int64_t)__ldg(seg_row + s0 + lp_i[r]) * L + cb0 + c);
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc.v[k] = FAST ? __fmaf_rn(wgt, v.v[k], acc.v[k])
                                    : __fadd_rn(acc.v[k], exact_div(__fmul_rn(wgt, v.v[k]), divisor, rdivisor));
            }
        }
        if (FAST) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] = exact_div(acc.v[k], divisor, rdivisor);
        }
        acc.store(out + (int64_t)PM[lp] * ldo + cb0 + c);
    }
}

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
static int tile_vec(int L, int CB, int64_t ld_a, const void *p0, const void *p1)
{
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    if (L % 4 == 0 && CB % 4 == 0 && ld_a % 4 == 0 && al(p0, 16) && al(p1, 16)) return 4;
    if (L % 2 == 0 && CB % 2 == 0 && ld_a % 2 == 0 && al(p0, 8) && al(p1, 8)) return 2;
    return 1;
}

static int check_tiles(const sgp_tiles_view *t, int L)
{
    if (!t || !t->perm) return fail(SGP_EINVAL, "null tiles view");
    if (t->N <= 0 || t->M <= 0 || t->S <= 0 || t->d < 1 || t->tile_points < 1 || L < 1)
        return fail(SGP_EINVAL, "bad tiles view");
    return SGP_OK;
}

extern "C" int sgp_splat_tiles(const sgp_tiles_view *t, const float *src, int64_t lds, int L, float *values,
                               sgp_stream_t stream)
{
    SGP_RANGE("sgp_splat_tiles");
    int rc = check_tiles(t, L);
    if (rc) return rc;
    if (!src || !values || lds < L || !t->tile_piece_ptr || !t->piece_ptr || !t->piece_row || !t->seg_ent)
        return fail(SGP_EINVAL, "sgp_splat_tiles: null pointer or lds < L");
    cudaStream_t st = (cudaStream_t)stream;
    const int CB = L < 16 ? L : 16;
    const int vec = tile_vec(L, CB, lds, src, values);
    const int T = t->tile_points, dp1 = t->d + 1;
    const int64_t n_tiles = (t->N + T - 1) / T;
    const unsigned ncb = (unsigned)((L + CB - 1) / CB);
    const size_t smem = (size_t)T * CB * 4 + (size_t)T * dp1 * 8 + ((size_t)T * dp1 + 8) * 4 + (size_t)T * dp1 * 4 + (size_t)T * 4;
    if (smem > 227 * 1024) return fail(SGP_EUNSUPPORTED, "splat tiles need %zu bytes of shared memory", smem);
    CUDA_TRY(cudaMemsetAsync(values, 0, sizeof(float) * (size_t)t->M * (size_t)L, st));
    dim3 grid((unsigned)n_tiles, ncb);
#define SGP_LAUNCH_SPLAT_TILES(VV)                                                                                   \
    do {                                                                                                             \
        if (smem > 48 * 1024) /* per (function, device): set on every launch, see sgp_groups.cu */                   \
            CUDA_TRY(cudaFuncSetAttribute(sgp_splat_tiles_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem));                                                               \
        sgp_splat_tiles_kernel<VV><<<grid, TILE_THREADS, smem, st>>>(t->perm, t->tile_piece_ptr, t->piece_ptr,      \
                                                                     t->piece_row, (const int2 *)t->seg_ent, src,    \
                                                                     lds, t->N, T, dp1, L, CB, values);              \
    } while (0)
    if (vec == 4) SGP_LAUNCH_SPLAT_TILES(4);
    else if (vec == 2) SGP_LAUNCH_SPLAT_TILES(2);
    else SGP_LAUNCH_SPLAT_TILES(1);
#undef SGP_LAUNCH_SPLAT_TILES
    return launch_ok("sgp_splat_tiles_kernel");
}

extern "C" int sgp_slice_tiles(const sgp_tiles_view *t, const float *values, int L, float *out, int64_t ldo,
                               int fast, sgp_stream_t stream)
{
    SGP_RANGE("sgp_slice_tiles");
    int rc = check_tiles(t, L);
    if (rc) return rc;
    if (!values || !out || ldo < L || !t->tile_seg_ptr || !t->seg_row || !t->lidx || !t->tile_w)
        return fail(SGP_EINVAL, "sgp_slice_tiles: null pointer or ldo < L");
    cudaStream_t st = (cudaStream_t)stream;
    const int CB = L < 16 ? L : 16;
    const int vec = tile_vec(L, CB, ldo, values, out);
    const int T = t->tile_points, dp1 = t->d + 1;
    const int64_t n_tiles = (t->N + T - 1) / T;
    const unsigned ncb = (unsigned)((L + CB - 1) / CB);
    // dictionary rows staged per CTA (tiles with more read the lattice directly)
    static int cap_env = 0;   // tuning hook: SGP_TILE_DICT_CAP
    if (cap_env == 0) {
        const char *e = getenv("SGP_TILE_DICT_CAP");
        cap_env = e ? atoi(e) : 768;
        if (cap_env < 1) cap_env = 768;
    }
    int cap = t->dict_cap < cap_env ? t->dict_cap : cap_env;
    if (cap < 1) cap = 1;
    const size_t smem = (((size_t)cap * CB + 3) & ~(size_t)3) * 4 + (size_t)T * dp1 * 4 + (size_t)((cap + 3) & ~3) * 4 + (size_t)T * 4 +
                        (((size_t)T * dp1 * 2 + 15) & ~(size_t)15);
    if (smem > 227 * 1024) return fail(SGP_EUNSUPPORTED, "slice tiles need %zu bytes of shared memory", smem);
    const float divisor = sgp_slice_divisor(t->d);
    volatile float rdivisor = 1.0f / divisor;
    dim3 grid((unsigned)n_tiles, ncb);
#define SGP_LAUNCH_SLICE_TILES(VV, FF)                                                                               \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            CUDA_TRY(cudaFuncSetAttribute(sgp_slice_tiles_kernel<VV, FF>,                                            \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        sgp_slice_tiles_kernel<VV, FF><<<grid, TILE_THREADS, smem, st>>>(t->perm, t->tile_seg_ptr, t->seg_row,      \
                                                                         t->lidx, t->tile_w, values, t->N, T, dp1,   \
                                                                         L, CB, cap, divisor, rdivisor, out, ldo);   \
    } while (0)
    if (fast) {
        if (vec == 4) SGP_LAUNCH_SLICE_TILES(4, true);
        else if (vec == 2) SGP_LAUNCH_SLICE_TILES(2, true);
        else SGP_LAUNCH_SLICE_TILES(1, true);
    } else {
        if (vec == 4) SGP_LAUNCH_SLICE_TILES(4, false);
        else if (vec == 2) SGP_LAUNCH_SLICE_TILES(2, false);
        else SGP_LAUNCH_SLICE_TILES(1, false);
    }
#undef SGP_LAUNCH_SLICE_TILES
    return launch_ok("sgp_slice_tiles_kernel");
}
