// sgp_groups.cu -- blur along several lattice axes per launch, staged through shared memory (B200, sm_100a).
//
// Reference semantics: PermutohedralLattice::blur, gpytorch_lattice_kernel/cpp/permutohedral.h:513-572 -- d+1
// sequential stencil passes, pass j combining each lattice point with its neighbours along direction j.
//
// Why groups.  In the basis a_i = (key_i - key_d)/(d+1) the lattice is Z^d: a step along axis j < d changes only
// a_j (by +-1), a step along axis d changes every a_i by -+1.  Hence for a set J of consecutive axes the lattice
// points fall into CLASSES -- points that agree on every coordinate difference a_i - a_k with i, k outside J -- and the
// passes of the axes in J never leave a class.  A CTA that holds whole classes in shared memory can therefore run all
// |J| passes of the group on chip: the lattice values are read once and written once per GROUP instead of being read
// 1+2r times and written once per AXIS (4x fewer L2 row transfers per axis at order 1 for groups of three).  At the
// metric configuration (N=1M, d=8, M=0.4M) classes of three-axis groups hold 31 points on average and 252 at most.
//
// Layout per group (built once per lattice by sgp_group_prepare / sgp_group_finalize):
//   order[p]        lattice point (first-touch index) stored at position p: points sorted by class
//   batch_begin[b]  CTA b owns positions [batch_begin[b], batch_begin[b+1]) -- whole classes, at most rows_cap rows
//   src[p]          where position p's input row lives in the PREVIOUS stage's output order (gather on load)
//   lnb             per batch [axis a][row][t]: position, relative to batch_begin[b], of the neighbour along the group's
//                   a-th axis, offset t over o = -r..-1, 1..r; absent = index of the kernel's all-zero row (512 or 1024)
// The arithmetic of a pass is that of sgp_blur_kernel (same products, same order, no FMA), so results are
// bit-identical to the per-axis path.
//
// Sorting / scanning in the build uses CUB (CUDA toolkit header library); the MVM kernel is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok
#define grid_for sgp_grid_for

#define LNB_ABSENT 0xFFFFu

// ------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix_u64(uint64_t h)
{
    h ^= h >> 33;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    h *= 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 33;
    return h;
}

// class hash of lattice point i for the axis range [j0, j1): hash of (a_c - a_k) over the axes c outside the range,
// k being the first axis outside it.  Points of one class get equal hashes; distinct classes that collide are merely
// processed together (harmless).
__global__ void __launch_bounds__(256)
sgp_group_hash_kernel(const int16_t *__restrict__ keys, int64_t M, int d, int j0, int j1, int shift,
                      unsigned long long *__restrict__ hash, uint32_t *__restrict__ ids)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int16_t *kp = keys + i * d;
    int sum = 0;
    for (int c = 0; c < d; ++c) sum += kp[c];
    // coordinate c of the full (d+1)-vector: stored for c < d, minus the sum for c = d
    auto coord = [&](int c) { return c < d ? (int)kp[c] : -sum; };
    int k = -1;
    for (int c = 0; c <= d; ++c)
        if (c < j0 || c >= j1) { k = c; break; }
    uint64_t h = 0x9E3779B97F4A7C15ull;
    if (k >= 0) {
        const int ck = coord(k);
        for (int c = k + 1; c <= d; ++c) {
            if (c >= j0 && c < j1) continue;
            const int diff = (coord(c) - ck) / (d + 1);   // exact: all coordinates share one remainder
            h = mix_u64(h ^ (uint64_t)(uint32_t)diff);
        }
    }
    hash[i] = h >> shift;     // the top bits only: fewer radix passes (see group_hash_bits)
    ids[i] = (uint32_t)i;
}

// head[p] = p if position p starts a class else 0; an inclusive max-scan turns it into class_start[p]
__global__ void __launch_bounds__(256)
sgp_group_heads_kernel(const unsigned long long *__restrict__ sorted_hash, int64_t M, uint32_t *__restrict__ head)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    head[p] = (p > 0 && sorted_hash[p] != sorted_hash[p - 1]) ? (uint32_t)p : 0u;
}

__global__ void __launch_bounds__(256)
sgp_group_pos_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ class_start, int64_t M,
                     uint32_t *__restrict__ pos, uint32_t *__restrict__ max_class)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    pos[order[p]] = (uint32_t)p;
    if (p == M - 1 || class_start[p + 1] != class_start[p]) atomicMax(max_class, (uint32_t)(p - class_start[p] + 1));
}

// Greedy packing of whole classes into CTA batches of at most `cap` rows: a batch starts where the previous one ended
// and extends to the last class boundary within cap rows, i.e. next = class_start[begin + cap].  The chain is
// sequential (each hop needs the previous one), so it is walked by one thread -- but out of shared memory: the block
// streams class_start through a window of PACK_WINDOW entries (coalesced loads by all threads), and the walker hops
// inside the window at shared-memory latency (~30 cycles a hop instead of a ~500 ns global round trip).
// out[0] = n_batches, out[1] = rows of the largest batch, out[2] = error flag.
#define PACK_WINDOW 25600      /* two windows of 100 KB of dynamic shared memory, used alternately */
#define PACK_THREADS 1024
// The walker (thread 0) reads class_start at the positions b + cap, which grow by at most cap <= 1024 per hop: the
// block streams class_start through consecutive windows [cap + k W, cap + (k+1) W); while the walker hops inside window k
// at shared-memory latency, all threads have the copy of window k+1 in flight (cp.async), so the global round trips of
// the stream are hidden behind the walk instead of alternating with it (146 -> ~60 us per group at M = 4e5).
__global__ void __launch_bounds__(PACK_THREADS)
sgp_group_pack_kernel(const uint32_t *__restrict__ class_start, int64_t M, int64_t cap, int64_t max_batches,
                      uint32_t *__restrict__ batch_begin, uint32_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t win[];   // 2 x PACK_WINDOW entries
    __shared__ long long s_begin, s_nb;
    __shared__ uint32_t s_max, s_err, s_done;
    if (threadIdx.x == 0) { s_begin = 0; s_nb = 0; s_max = 0; s_err = 0; s_done = 0; }
    const int64_t n_win = M > cap ? (M - cap + PACK_WINDOW - 1) / PACK_WINDOW : 0;
    auto load = [&](int64_t k, uint32_t *dst) {     // window k = class_start[cap + k W, cap + (k+1) W) clipped to M
        const int64_t w0 = cap + k * PACK_WINDOW;
        const int64_t wn = (M - w0 < PACK_WINDOW) ? M - w0 : PACK_WINDOW;
        if (wn > 0) cta_copy_async(dst, class_start + w0, (int)(wn * 4), threadIdx.x, PACK_THREADS);
    };
    if (n_win > 0) load(0, win);
    cp_async_wait_all();
    __syncthreads();
    for (int64_t k = 0; k <= n_win; ++k) {
        uint32_t *cur = win + (k & 1) * PACK_WINDOW;
        if (k + 1 < n_win) load(k + 1, win + ((k + 1) & 1) * PACK_WINDOW);
        if (threadIdx.x == 0 && !s_done) {
            const int64_t w0 = cap + k * PACK_WINDOW;
            const int64_t w1 = (k < n_win) ? ((w0 + PACK_WINDOW < M) ? w0 + PACK_WINDOW : M) : M;   // window k covers [w0, w1)
            int64_t b = s_begin, nb = s_nb;
            uint32_t mx = s_max;
            while (b < M && nb < max_batches) {
                int64_t end = b + cap;
                if (end >= M) {
                    end = M;
                } else {
                    if (end >= w1) break;                  // the next read lies in the next window
                    end = cur[end - w0];                   // start of the class that position b + cap falls into
                    if (end <= b) { s_err = 1; break; }    // a class larger than cap (the caller checks max_class)
                }
                batch_begin[nb++] = (uint32_t)b;
                if ((uint32_t)(end - b) > mx) mx = (uint32_t)(end - b);
                b = end;
            }
            s_begin = b; s_nb = nb; s_max = mx;
            if (b >= M || nb >= max_batches || s_err) s_done = 1;
        }
        cp_async_wait_all();
        __syncthreads();
        if (s_done) break;
    }
    if (threadIdx.x == 0) {
        if (s_begin < M) s_err = 1;
        batch_begin[s_nb] = (uint32_t)M;
        out[0] = (uint32_t)s_nb;
        out[1] = s_max;
        out[2] = s_err;
    }
}

// Neighbours come from the whole-lattice table nbr when it exists, else straight from the key hash table (the lattice
// build then never materialises the (d+1) x M x 2r neighbour table: 150 GB at the stress configuration).
__global__ void __launch_bounds__(256)
sgp_group_tables_kernel(const int32_t *__restrict__ nbr, const int16_t *__restrict__ keys, int d,
                        const unsigned long long *__restrict__ table, uint64_t mask, int64_t M, int order_r, int j0, int j1,
                        const uint32_t *__restrict__ order, const uint32_t *__restrict__ pos,
                        const uint32_t *__restrict__ prev_pos, const uint32_t *__restrict__ n_batches_dev,
                        const uint32_t *__restrict__ batch_begin, uint32_t absent, int32_t *__restrict__ src,
                        uint16_t *__restrict__ lnb, int32_t *__restrict__ flags)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const int64_t n_batches = (int64_t)__ldg(n_batches_dev);   // left by the pack kernel: no host round trip in between
    if (n_batches < 1) return;                                  // (the pack failed: its error flag says so)
    const uint32_t row = order[p];
    src[p] = prev_pos ? (int32_t)prev_pos[row] : (int32_t)row;
    // batch of position p: last b with batch_begin[b] <= p
    int64_t lo = 0, hi = n_batches;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)batch_begin[mid] <= p) lo = mid; else hi = mid;
    }
    const uint32_t base = batch_begin[lo], limit = batch_begin[lo + 1];
    const int nax = j1 - j0, w = 2 * order_r;
    for (int a = 0; a < nax; ++a) {
        const int j = j0 + a;
        for (int t = 0; t < w; ++t) {
            int32_t nb;
            if (nbr) {
                nb = nbr[((int64_t)j * M + row) * w + t];
            } else {   // key - o on every stored coordinate, key[j] + o*d on coordinate j (permutohedral.h:541-542)
                const int o = (t < order_r) ? t - order_r : t - order_r + 1;
                const int16_t *kp = keys + (int64_t)row * d;
                int16_t nk[SGP_MAX_DIM];
                for (int c = 0; c < d; ++c) nk[c] = (int16_t)((int)kp[c] - o);
                if (j < d) nk[j] = (int16_t)((int)kp[j] + o * d);
                nb = sgp_table_find(keys, table, mask, nk, d);
            }
            uint32_t local = absent;                   // index of the kernel's all-zero row
            if (nb >= 0) {
                const uint32_t q = pos[nb];
                local = q - base;
                if (q < base || q >= limit || local >= absent) {   // neighbour outside the CTA's rows: must not happen
                    atomicOr(flags, 4);
                    local = absent;
                }
            }
            // per batch: [axis][row][t], so that a thread's rows of one pass are a fixed stride apart
            lnb[(int64_t)base * nax * w + ((int64_t)a * (limit - base) + (p - base)) * w + t] = (uint16_t)local;
        }
    }
}

// replay_out[pv] = {pos[replay[pv].index], replay[pv].weight}: the slice reads the last stage's output order
__global__ void __launch_bounds__(256)
sgp_remap_replay_kernel(const int2 *__restrict__ replay, int64_t total, const uint32_t *__restrict__ pos,
                        int2 *__restrict__ out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const int2 e = replay[q];
    out[q] = make_int2((int)pos[e.x], e.y);
}

struct GroupWs {
    size_t hash_a, hash_b, ids_a, head, cub, cub_bytes, small, bytes;
};

static int group_ws_layout(int64_t M, GroupWs *w)
{
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    w->hash_a = o; o += al((size_t)M * 8);
    w->hash_b = o; o += al((size_t)M * 8);
    w->ids_a = o; o += al((size_t)M * 4);
    w->head = o; o += al((size_t)M * 4);
    w->small = o; o += 256;
    size_t t1 = 0, t2 = 0;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long *)nullptr,
                                                     (unsigned long long *)nullptr, (const uint32_t *)nullptr,
                                                     (uint32_t *)nullptr, (int64_t)M, 0, 64);
    cudaError_t e2 = cub::DeviceScan::InclusiveScan(nullptr, t2, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                    cub::Max(), (int64_t)M);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return sgp_fail(SGP_ECUDA, "cub temp-size query failed");
    w->cub_bytes = t1 > t2 ? t1 : t2;
    w->cub = o; o += al(w->cub_bytes);
    w->bytes = o;
    return SGP_OK;
}

// Bits of class hash that are sorted: log2(M) + 20, rounded up to whole 8-bit radix passes.  Two of at most M classes
// then collide with probability below 2^-20 per class -- about one merged pair per million classes, and a merged
// pair is merely processed together.  40 bits (5 passes instead of 8) at M = 4e5, 48 at M = 2.5e8.
static int group_hash_bits(int64_t M)
{
    int lg = 0;
    while ((1ll << lg) < M) ++lg;
    int bits = (lg + 20 + 7) / 8 * 8;
    return bits > 64 ? 64 : bits;
}

extern "C" size_t sgp_group_workspace_bytes(int64_t M)
{
    GroupWs w;
    if (M <= 0 || group_ws_layout(M, &w) != SGP_OK) return 0;
    return w.bytes;
}

// result (device uint32[8], one per group): [0] largest class, [1] batches, [2] rows of the largest batch, [3] pack error,
// [4] neighbour-outside-batch flag.  The *_async forms launch everything on the stream and never synchronise: a caller
// that knows the axis ranges (from the last lattice of the same shape) builds all groups back to back and reads all the
// results with ONE synchronisation; sgp_group_prepare / sgp_group_finalize are the same launches plus the read-back.
static int group_prepare_launch(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                                uint32_t *class_start, void *workspace, size_t workspace_bytes, uint32_t *result,
                                cudaStream_t st)
{
    if (!keys || !order || !pos || !class_start || !workspace || !result || M <= 0 || d < 1 || d > SGP_MAX_DIM || j0 < 0 ||
        j1 <= j0 || j1 > d + 1)
        return fail(SGP_EINVAL, "sgp_group_prepare: bad argument");
    if (M >= (1ll << 32)) return fail(SGP_EOVERFLOW, "M does not fit 32 bits");
    GroupWs w;
    int rc = group_ws_layout(M, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "group workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    char *base = (char *)workspace;
    unsigned long long *ha = (unsigned long long *)(base + w.hash_a), *hb = (unsigned long long *)(base + w.hash_b);
    uint32_t *ids = (uint32_t *)(base + w.ids_a), *head = (uint32_t *)(base + w.head);
    size_t cub_bytes = w.cub_bytes;
    const int bits = group_hash_bits(M);
    sgp_group_hash_kernel<<<grid_for(M, 256), 256, 0, st>>>(keys, M, d, j0, j1, 64 - bits, ha, ids);
    rc = launch_ok("sgp_group_hash_kernel");
    if (rc) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(base + w.cub, cub_bytes, ha, hb, ids, order, (int64_t)M, 0, bits, st));
    sgp_group_heads_kernel<<<grid_for(M, 256), 256, 0, st>>>(hb, M, head);
    rc = launch_ok("sgp_group_heads_kernel");
    if (rc) return rc;
    cub_bytes = w.cub_bytes;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(base + w.cub, cub_bytes, head, class_start, cub::Max(), (int64_t)M, st));
    CUDA_TRY(cudaMemsetAsync(result, 0, 32, st));
    sgp_group_pos_kernel<<<grid_for(M, 256), 256, 0, st>>>(order, class_start, M, pos, result);
    return launch_ok("sgp_group_pos_kernel");
}

static int group_finalize_launch(const int32_t *nbr, const int16_t *keys, int d, const uint64_t *table, int64_t capacity,
                                 int64_t M, int order_r, int j0, int j1, const uint32_t *order, const uint32_t *pos,
                                 const uint32_t *class_start, const uint32_t *prev_pos, int64_t cap, int64_t max_batches,
                                 uint32_t *batch_begin, int32_t *src, uint16_t *lnb, uint32_t *result, cudaStream_t st)
{
    if (!nbr && (!keys || !table || capacity < 2 || (capacity & (capacity - 1)) != 0 || d < 1 || d > SGP_MAX_DIM))
        return fail(SGP_EINVAL, "sgp_group_finalize: needs either nbr or keys + hash table");
    if (!order || !pos || !class_start || !batch_begin || !src || !lnb || !result || M <= 0 || order_r < 1 ||
        order_r > SGP_MAX_ORDER || j1 <= j0 || cap < 1 || cap > 1024 || max_batches < 1)
        return fail(SGP_EINVAL, "sgp_group_finalize: bad argument");
    // the opt-in above 48 KB is per device (and cheap): set it on every call rather than caching it per process
    CUDA_TRY(cudaFuncSetAttribute(sgp_group_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * PACK_WINDOW * sizeof(uint32_t))));
    sgp_group_pack_kernel<<<1, PACK_THREADS, 2 * PACK_WINDOW * sizeof(uint32_t), st>>>(class_start, M, cap, max_batches,
                                                                                       batch_begin, result + 1);
    int rc = launch_ok("sgp_group_pack_kernel");
    if (rc) return rc;
    const uint32_t absent = cap > 512 ? 1024u : 512u;   // ROWS_MAX of the kernel variant that takes this group
    sgp_group_tables_kernel<<<grid_for(M, 256), 256, 0, st>>>(nbr, keys, d, (const unsigned long long *)table,
                                                               (uint64_t)(capacity - 1), M, order_r, j0, j1, order, pos, prev_pos,
                                                               result + 1, batch_begin, absent, src, lnb, (int32_t *)(result + 4));
    return launch_ok("sgp_group_tables_kernel");
}

extern "C" int sgp_group_prepare_async(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                                       uint32_t *class_start, void *workspace, size_t workspace_bytes, uint32_t *result,
                                       sgp_stream_t stream)
{
    SGP_RANGE("sgp_group_prepare_async");
    return group_prepare_launch(keys, M, d, j0, j1, order, pos, class_start, workspace, workspace_bytes, result,
                                (cudaStream_t)stream);
}

extern "C" int sgp_group_finalize_async(const int32_t *nbr, const int16_t *keys, int d, const uint64_t *table,
                                        int64_t capacity, int64_t M, int order_r, int j0, int j1, const uint32_t *order,
                                        const uint32_t *pos, const uint32_t *class_start, const uint32_t *prev_pos,
                                        int64_t cap, int64_t max_batches, uint32_t *batch_begin, int32_t *src, uint16_t *lnb,
                                        uint32_t *result, sgp_stream_t stream)
{
    SGP_RANGE("sgp_group_finalize_async");
    return group_finalize_launch(nbr, keys, d, table, capacity, M, order_r, j0, j1, order, pos, class_start, prev_pos, cap,
                                 max_batches, batch_begin, src, lnb, result, (cudaStream_t)stream);
}

extern "C" int sgp_group_prepare(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                                 uint32_t *class_start, void *workspace, size_t workspace_bytes, int64_t *max_class_out,
                                 sgp_stream_t stream)
{
    SGP_RANGE("sgp_group_prepare");
    if (!max_class_out) return fail(SGP_EINVAL, "sgp_group_prepare: bad argument");
    GroupWs w;
    int rc = M > 0 ? group_ws_layout(M, &w) : fail(SGP_EINVAL, "sgp_group_prepare: bad argument");
    if (rc) return rc;
    if (!workspace || workspace_bytes < w.bytes) return fail(SGP_EINVAL, "group workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *small = (uint32_t *)((char *)workspace + w.small);
    rc = group_prepare_launch(keys, M, d, j0, j1, order, pos, class_start, workspace, workspace_bytes, small, st);
    if (rc) return rc;
    uint32_t mx = 0;
    CUDA_TRY(cudaMemcpyAsync(&mx, small, sizeof(mx), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *max_class_out = (int64_t)mx;
    return SGP_OK;
}

extern "C" int64_t sgp_group_max_batches(int64_t M, int64_t cap, int64_t max_class)
{
    if (M <= 0 || cap < 1 || max_class > cap) return 0;
    return M / (cap - max_class + 1) + 2;   // every batch but the last holds more than cap - max_class rows
}

extern "C" int sgp_group_finalize(const int32_t *nbr, const int16_t *keys, int d, const uint64_t *table, int64_t capacity,
                                  int64_t M, int order_r, int j0, int j1, const uint32_t *order,
                                  const uint32_t *pos, const uint32_t *class_start, const uint32_t *prev_pos,
                                  int64_t cap, int64_t max_batches, uint32_t *batch_begin, int32_t *src, uint16_t *lnb,
                                  void *workspace, size_t workspace_bytes, int64_t *n_batches_out, int32_t *max_rows_out,
                                  sgp_stream_t stream)
{
    SGP_RANGE("sgp_group_finalize");
    if (!workspace || !max_rows_out || !n_batches_out || M <= 0) return fail(SGP_EINVAL, "sgp_group_finalize: bad argument");
    GroupWs w;
    int rc = group_ws_layout(M, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "group workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *small = (uint32_t *)((char *)workspace + w.small);
    CUDA_TRY(cudaMemsetAsync(small, 0, 32, st));
    rc = group_finalize_launch(nbr, keys, d, table, capacity, M, order_r, j0, j1, order, pos, class_start, prev_pos, cap,
                               max_batches, batch_begin, src, lnb, small, st);
    if (rc) return rc;
    uint32_t host[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(host, small, sizeof(host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (host[3] != 0) return fail(SGP_EINVAL, "blur group [%d,%d): classes do not fit %lld rows", j0, j1, (long long)cap);
    if (host[4] != 0) return fail(SGP_EINVAL, "blur group [%d,%d): a neighbour fell outside its CTA batch", j0, j1);
    *n_batches_out = host[1];
    *max_rows_out = (int32_t)host[2];
    return SGP_OK;
}

extern "C" int sgp_remap_replay(const int32_t *replay, int64_t total, const uint32_t *pos, int32_t *replay_out,
                                sgp_stream_t stream)
{
    if (total == 0) return SGP_OK;
    if (!replay || !pos || !replay_out || total < 0) return fail(SGP_EINVAL, "sgp_remap_replay: bad argument");
    sgp_remap_replay_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const int2 *)replay, total, pos,
                                                                                     (int2 *)replay_out);
    return launch_ok("sgp_remap_replay_kernel");
}

// ------------------------------------------------------------------------------------
// the MVM kernel: one CTA per batch (x one block of CHUNKS*VEC channels)
// ------------------------------------------------------------------------------------
struct GroupCoeffs {
    float c[2 * SGP_MAX_ORDER + 1];
};

#define PASS_UNROLL 4

// Shared memory: A[ROWS_MAX+1][CBT] | nbs[nax][rows][2r] (uint16) | SRC[rows]; CBT = CHUNKS*VEC channels,
// ROWS_MAX = 2*THREADS rows (the most a CTA of this variant takes), row ROWS_MAX is all zeros.
//
// A thread owns one channel chunk (c) and the rows lr0 + i*RSTEP (i < MAXI) of the batch for the whole kernel:
//   * the index arithmetic is hoisted out of every loop (row slots are compile-time offsets);
//   * a row's own value never has to be read back from shared memory -- it stays in registers from the pass that
//     produced it; shared memory holds ONE value buffer that a pass reads (neighbours), then, after a barrier, every
//     thread overwrites its rows in place;
//   * an absent neighbour is the index ROWS_MAX of the zero row (written by the builder), exactly what the reference
//     does (permutohedral.h:545 substitutes a zero vector and still adds c*0): the passes are branch-free;
//   * PASS_UNROLL rows are processed together -- their neighbour indices, then all their neighbour rows, are read
//     before any arithmetic -- so that several shared-memory round trips overlap (warps issue in order).
template <int VEC, int R, int CHUNKS, int THREADS, bool FAST, int ROWS_MAX_T = 2 * THREADS>
__global__ void __launch_bounds__(THREADS)   // (capping registers for a fourth CTA per SM measured slower: 62 vs 56 us)
sgp_blur_group_kernel(const uint32_t *__restrict__ batch_begin, const int32_t *__restrict__ src,
                      const uint16_t *__restrict__ lnb, const float *__restrict__ in, float *__restrict__ out, int L,
                      int rows_cap, int nax, int order_rt, GroupCoeffs cf)
{
    constexpr int RR = R > 0 ? R : SGP_MAX_ORDER;
    constexpr int CBT = CHUNKS * VEC;
    constexpr int RSTEP = THREADS / CHUNKS;
    constexpr int U = PASS_UNROLL;
    constexpr int ROWS_MAX = ROWS_MAX_T;                     // 512 rows at 256 threads, 1024 at 512 (host-checked); (512 rows
                                                             // on 128 threads for one-chunk rows measured no faster: the
                                                             // narrow-row blur is bound by its neighbour tables)
    constexpr int MAXI = ROWS_MAX / RSTEP;                   // row slots per thread: 2 * CHUNKS
    constexpr int UU = MAXI < U ? MAXI : U;
    static_assert(MAXI % UU == 0, "row slots must split into whole unroll groups");
    const int r = R > 0 ? R : order_rt;
    const int w2 = 2 * r;
    extern __shared__ __align__(16) float smem[];
    float *A = smem;
    uint16_t *nbs = (uint16_t *)(smem + (((ROWS_MAX + 1) * CBT + 3) & ~3));   // 16-byte aligned (cp.async)
    const uint32_t p0 = batch_begin[blockIdx.x];
    const int rows = (int)(batch_begin[blockIdx.x + 1] - p0);
    if (rows == 0) return;
    const int c = (threadIdx.x % CHUNKS) * VEC;
    const int lr0 = threadIdx.x / CHUNKS;
    const int cg = blockIdx.y * CBT + c;          // first global channel of this thread
    const bool live = cg < L;                      // the last channel block may be partial (L % CBT != 0)

    pdl_launch_dependents();
    if (threadIdx.x < CBT) A[ROWS_MAX * CBT + threadIdx.x] = 0.0f;
    // stage the batch's neighbour table and its slice of the gather list (plain range copies, asynchronous) ...
    int32_t *SRC = (int32_t *)(nbs + (((size_t)rows_cap * nax * w2 + 7) & ~(size_t)7));
    cta_copy_async(nbs, lnb + (int64_t)p0 * nax * w2, rows * nax * w2 * 2, threadIdx.x, THREADS);   // w2 even: whole words
    cta_copy_async(SRC, src + p0, rows * 4, threadIdx.x, THREADS);
    cp_async_wait_all();
    __syncthreads();
    // ... and the rows themselves, gathered from the previous stage's order, all in flight at once (cp.async).
    // Everything above read build-time tables only; the rows are the previous kernel's output: wait for it here.
    pdl_wait();
    if (live) {
        for (int lr = lr0; lr < rows; lr += RSTEP) cp_async_vec<VEC>(A + lr * CBT + c, in + (int64_t)SRC[lr] * L + cg);
    }
    cp_async_wait_all();
    __syncthreads();
    // (threads of a partial last channel block carry on with scratch values: they keep the barriers uniform and never store)

    float *const Ac = A + c;                                 // this thread's channel chunk of row 0
    Vec<VEC> self[MAXI];
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
        const int lr = lr0 + i * RSTEP;
        self[i].load_plain(Ac + (lr < rows ? lr : ROWS_MAX) * CBT);
    }
    for (int a = 0; a < nax; ++a) {
        const bool last = (a == nax - 1);
        const uint16_t *nb_a = nbs + (size_t)a * rows * w2 + lr0 * w2;   // [a][lr][w2], this thread's first row
#pragma unroll
        for (int i0 = 0; i0 < MAXI; i0 += UU) {
            if (lr0 + i0 * RSTEP < rows) {
                uint32_t ni[UU][2 * RR];
#pragma unroll
                for (int u = 0; u < UU; ++u) {
                    const int lr = lr0 + (i0 + u) * RSTEP;
                    const uint16_t *nb = nb_a + (i0 + u) * RSTEP * w2;
                    if (R == 1) {
                        const uint32_t both = (lr < rows) ? *(const uint32_t *)nb : (uint32_t)(ROWS_MAX | (ROWS_MAX << 16));
                        ni[u][0] = both & 0xFFFFu;
                        ni[u][1] = both >> 16;
                    } else {
#pragma unroll
                        for (int t = 0; t < 2 * RR; ++t) ni[u][t] = (t < w2 && lr < rows) ? (uint32_t)nb[t] : (uint32_t)ROWS_MAX;
                    }
                }
                Vec<VEC> acc[UU];
#pragma unroll
                for (int u = 0; u < UU; ++u) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[u].v[k] = 0.0f;
                }
                // reference order: o = -r..-1, 0, 1..r
#pragma unroll
                for (int t = 0; t < RR; ++t) {
                    if (t < r) {
#pragma unroll
                        for (int u = 0; u < UU; ++u) {
                            Vec<VEC> v;
                            v.load_plain(Ac + ni[u][t] * CBT);
#pragma unroll
                            for (int k = 0; k < VEC; ++k) acc[u].v[k] = madd<FAST>(cf.c[t], v.v[k], acc[u].v[k]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UU; ++u) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[u].v[k] = madd<FAST>(cf.c[r], self[i0 + u].v[k], acc[u].v[k]);
                }
#pragma unroll
                for (int t = 0; t < RR; ++t) {
                    if (t < r) {
#pragma unroll
                        for (int u = 0; u < UU; ++u) {
                            Vec<VEC> v;
                            v.load_plain(Ac + ni[u][r + t] * CBT);
#pragma unroll
                            for (int k = 0; k < VEC; ++k) acc[u].v[k] = madd<FAST>(cf.c[r + 1 + t], v.v[k], acc[u].v[k]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UU; ++u) {
                    const int lr = lr0 + (i0 + u) * RSTEP;
                    self[i0 + u] = acc[u];
                    if (last && live && lr < rows) acc[u].store(out + (int64_t)(p0 + lr) * L + cg);   // the group's result
                }
            }
        }
        if (!last) {
            __syncthreads();   // every neighbour read of this pass is done: the rows can be overwritten in place
#pragma unroll
            for (int i = 0; i < MAXI; ++i) self[i].store(Ac + (lr0 + i * RSTEP) * CBT);   // rows >= `rows`: scratch
            __syncthreads();
        }
    }
}

static size_t group_smem_bytes(int rows_cap, int rows_max, int cbt, int nax, int order)
{
    return ((((size_t)rows_max + 1) * (size_t)cbt + 3) & ~(size_t)3) * sizeof(float) +
           (((size_t)rows_cap * nax * 2 * order * sizeof(uint16_t) + 15) & ~(size_t)15) + (size_t)rows_cap * sizeof(int32_t);
}

template <int VEC, int CHUNKS, int THREADS, bool FAST, int ROWS_MAX_T = 2 * THREADS>
static int launch_group(const sgp_blur_group *g, int order, const GroupCoeffs &cf, int L, const float *in, float *out,
                        cudaStream_t st)
{
    constexpr int CBT = VEC * CHUNKS;
    const int nax = g->j1 - g->j0;
    const size_t smem = group_smem_bytes(g->rows_cap, ROWS_MAX_T, CBT, nax, order);
    if (smem > 227 * 1024)
        return fail(SGP_EUNSUPPORTED, "blur group needs %zu bytes of shared memory (%d rows x %d channels)", smem,
                    g->rows_cap, CBT);
    dim3 grid((unsigned)g->n_batches, (unsigned)((L + CBT - 1) / CBT));
#define SGP_LAUNCH_GROUP(RR)                                                                                          \
    do {                                                                                                              \
        /* the opt-in above 48 KB is an attribute of (function, device): set it on every launch (a host-side call of  \
           about a microsecond) instead of caching it per process, so that a second device works too */              \
        if (smem > 48 * 1024)                                                                                         \
            CUDA_TRY(cudaFuncSetAttribute(sgp_blur_group_kernel<VEC, RR, CHUNKS, THREADS, FAST, ROWS_MAX_T>,                \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        cudaError_t le = sgp_launch_pdl(sgp_blur_group_kernel<VEC, RR, CHUNKS, THREADS, FAST, ROWS_MAX_T>, grid, dim3(THREADS), smem, st, \
                                        g->batch_begin, g->src, g->lnb, in, out, L, g->rows_cap, nax, order, cf);     \
        if (le != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_blur_group_kernel failed: %s", cudaGetErrorString(le)); \
    } while (0)
    if (order == 1) SGP_LAUNCH_GROUP(1);
    else if (order == 2) SGP_LAUNCH_GROUP(2);
    else if (order == 3) SGP_LAUNCH_GROUP(3);
    else SGP_LAUNCH_GROUP(0);
#undef SGP_LAUNCH_GROUP
    return launch_ok("sgp_blur_group_kernel");
}

// channels staged per CTA: whole 64-byte row pieces when the rows are that wide
extern "C" int sgp_blur_groups_channel_block(int L)
{
    static int cb_env = -1;   // tuning hook: SGP_GROUP_CB=8 stages 32-byte row pieces (more CTAs per SM)
    if (cb_env < 0) {
        const char *e = getenv("SGP_GROUP_CB");
        cb_env = e ? atoi(e) : 0;
    }
    if (L % 4 == 0) {
        if (cb_env == 8 && L >= 8) return 8;
        return L >= 12 ? 16 : (L >= 8 ? 8 : 4);   // L = 12: one block of 16 with an idle quarter beats 8 + 4
    }
    if (L % 2 == 0) return L >= 8 ? 8 : (L >= 4 ? 4 : 2);
    return L >= 4 ? 4 : (L >= 2 ? 2 : 1);
}

extern "C" int sgp_blur_groups(const sgp_blur_group *groups, int n_groups, int64_t M, int order, const float *coeffs,
                               int k, int L, float *buf0, float *buf1, int *result_in_buf1, int fast, sgp_stream_t stream)
{
    SGP_RANGE("sgp_blur_groups");
    if (!groups || n_groups < 1 || M < 0 || L < 1 || !coeffs || k != 2 * order + 1 || order < 1 || order > SGP_MAX_ORDER)
        return fail(SGP_EINVAL, "sgp_blur_groups: bad argument");
    if (result_in_buf1) *result_in_buf1 = 0;
    if (M == 0) return SGP_OK;
    if (!buf0 || !buf1) return fail(SGP_EINVAL, "sgp_blur_groups: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    GroupCoeffs cf;
    memset(&cf, 0, sizeof(cf));
    memcpy(cf.c, coeffs, sizeof(float) * k);
    const int CB = sgp_blur_groups_channel_block(L);
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    int vec = 1;
    if (L % 4 == 0 && al(buf0, 16) && al(buf1, 16)) vec = 4;
    else if (L % 2 == 0 && al(buf0, 8) && al(buf1, 8)) vec = 2;
    const int chunks = CB / vec;   // 1, 2 or 4 by construction of CB
    float *in = buf0, *out = buf1;
    for (int gi = 0; gi < n_groups; ++gi) {
        const sgp_blur_group *g = groups + gi;
        if (!g->batch_begin || !g->src || !g->lnb || g->rows_cap < 1 || g->n_batches < 1 || g->j1 <= g->j0)
            return fail(SGP_EINVAL, "sgp_blur_groups: group %d is not built", gi);
        if (g->rows_cap > 1024) return fail(SGP_EUNSUPPORTED, "blur group %d: %d rows per CTA (limit 1024)", gi, g->rows_cap);
        if (g->zero_row != 512 && g->zero_row != 1024) return fail(SGP_EINVAL, "blur group %d: zero_row must be 512 or 1024", gi);
        if (g->rows_cap > g->zero_row) return fail(SGP_EINVAL, "blur group %d: %d rows exceed its variant's %d", gi, g->rows_cap, g->zero_row);
        const bool big = g->zero_row == 1024;   // 256 threads hold up to 512 rows, 512 threads up to 1024
        int rc = SGP_EUNSUPPORTED;
#define SGP_GROUP_CASE(VV, CC)                                                                          \
    if (vec == VV && chunks == CC)                                                                      \
        rc = fast ? (big ? launch_group<VV, CC, 512, true>(g, order, cf, L, in, out, st)   \
                         : launch_group<VV, CC, 256, true>(g, order, cf, L, in, out, st))  \
                  : (big ? launch_group<VV, CC, 512, false>(g, order, cf, L, in, out, st)  \
                         : launch_group<VV, CC, 256, false>(g, order, cf, L, in, out, st))
        SGP_GROUP_CASE(4, 4);
        else SGP_GROUP_CASE(4, 2);
        else SGP_GROUP_CASE(4, 1);
        else SGP_GROUP_CASE(2, 4);
        else SGP_GROUP_CASE(2, 2);
        else SGP_GROUP_CASE(2, 1);
        else SGP_GROUP_CASE(1, 4);
        else SGP_GROUP_CASE(1, 2);
        else SGP_GROUP_CASE(1, 1);
        else return fail(SGP_EUNSUPPORTED, "sgp_blur_groups: no kernel for vec=%d chunks=%d", vec, chunks);
#undef SGP_GROUP_CASE
        if (rc) return rc;
        float *t = in; in = out; out = t;
    }
    if (result_in_buf1) *result_in_buf1 = (in == buf1) ? 1 : 0;
    return SGP_OK;
}

// The production chain in one call: row-sorted splat -> blur-group stages -> slice.  `slice_view->replay` must address
// the lattice values in the order the last stage leaves them (sgp_permute_replay with that stage's pos).
// flags: SGP_MVM_PREZEROED (1) buf0 holds zeros on entry (skip the memset in front of the splat);
//        SGP_MVM_ZERO_AFTER (2) leave buf0 zeroed on exit.  With an odd number of stages buf0 is last read by the last
//        stage, so it is zeroed on a side stream WHILE the slice runs (fork / join with events: also inside a stream
//        capture, where it becomes a parallel branch of the graph); with an even number the slice itself reads buf0 and
//        the zeroing follows it.  Replaying a graph captured with both flags keeps the 25.6 MB memset of the metric
//        shape off the critical path of every product.
//        SGP_MVM_SRC_PADDED (4) src has L rounded up to a multiple of 4 columns (the caller's ragged block copied into a
//        zero-padded one).
struct CgEpilogue {          // sgp_mvm_rows_groups_cg: the sweep after the product, folded into the slice where possible
    const float *s, *noise;
    float *pAp, *scratch;
};

static int mvm_rows_groups_impl(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src,
                                int64_t lds, int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0,
                                float *buf1, int Lv, int flags, const CgEpilogue *cg, sgp_stream_t stream);

extern "C" int sgp_mvm_rows_groups_ex(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                      int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src,
                                      int64_t lds, int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0,
                                      float *buf1, int Lv, int flags, sgp_stream_t stream)
{
    return mvm_rows_groups_impl(slice_view, ent, seg_row, n_entries, groups, n_groups, src, lds, L, coeffs, k, out, ldo, buf0,
                                buf1, Lv, flags, nullptr, stream);
}

// out = s * K src + noise * src and pAp[l] = sum_n src[n, l] * out[n, l]: one CG iteration's product and the sweep that
// follows it (sgp_cg_apply), the sweep folded into the slice's epilogue when the TMA-ring slice applies (else it runs as
// its own launch).  out and src are [N, L] blocks with ld = L (what sgp_cg_apply expects); s, noise: device scalars;
// scratch: sgp_cg_scratch_floats(L) floats.
extern "C" int sgp_mvm_rows_groups_cg(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                      int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src,
                                      int64_t lds, int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0,
                                      float *buf1, int Lv, int flags, const float *s, const float *noise, float *pAp,
                                      float *scratch, sgp_stream_t stream)
{
    if (!s || !noise || !pAp || !scratch) return fail(SGP_EINVAL, "sgp_mvm_rows_groups_cg: null pointer");
    if (lds != L || ldo != L || (flags & 4) || (Lv != L && (L % 4 != 0 || Lv % 4 != 0 || Lv < L)))
        return fail(SGP_EINVAL, "sgp_mvm_rows_groups_cg: needs contiguous [N, L] blocks (lds = ldo = L) and Lv = L or, for L % 4 == 0, a multiple of 4 above it");
    const CgEpilogue cg{s, noise, pAp, scratch};
    return mvm_rows_groups_impl(slice_view, ent, seg_row, n_entries, groups, n_groups, src, lds, L, coeffs, k, out, ldo, buf0,
                                buf1, Lv, flags, &cg, stream);
}

// One whole CG iteration on the production chain, enqueued by ONE call (a Python loop that issues the product and the
// three sweeps one by one costs ~500 us of host time per iteration at N = 1M -- more than the 290 us the device needs):
//   AP = s K P + noise P, pAp            (sgp_mvm_rows_groups_cg)
//   alpha = rs / pAp; R -= alpha AP; rs_new, beta, done[it]      (sgp_cg_update_r; done[it] also copied to done_host[it])
//   X += alpha P; P = R + beta P         (sgp_cg_direction_x)
// alphas / betas / done: device [max_iter, L] / [max_iter, L] / [max_iter]; done_host: pinned host [max_iter] or NULL.
// The caller synchronises on an event recorded after the call before it reads done_host[it].
extern "C" int sgp_cg_iteration(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *coeffs, int k,
                                float *buf0, float *buf1, int flags, float *X, float *R, float *P, float *AP, float *rs,
                                float *pAp, const float *bnorm, const float *s, const float *noise, float tol,
                                int criterion, int L, int Lv, float *alphas, float *betas, int32_t *done,
                                int32_t *done_host, int it, float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_iteration");
    if (!slice_view || !X || !R || !P || !AP || !alphas || !betas || !done || it < 0)
        return fail(SGP_EINVAL, "sgp_cg_iteration: bad argument");
    const int64_t N = slice_view->N;
    int rc = sgp_mvm_rows_groups_cg(slice_view, ent, seg_row, n_entries, groups, n_groups, P, L, L, coeffs, k, AP, L, buf0, buf1,
                                    Lv, flags, s, noise, pAp, scratch, stream);
    if (rc) return rc;
    float *alpha = alphas + (size_t)it * L, *beta = betas + (size_t)it * L;
    rc = sgp_cg_update_r(R, AP, rs, pAp, bnorm, tol, criterion, N, L, alpha, beta, done + it, scratch, stream);
    if (rc) return rc;
    if (done_host)
        CUDA_TRY(cudaMemcpyAsync(done_host + it, done + it, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return sgp_cg_direction_x(P, R, X, alpha, beta, N, L, stream);
}

static int mvm_rows_groups_impl(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src,
                                int64_t lds, int L, const float *coeffs, int k, float *out, int64_t ldo, float *buf0,
                                float *buf1, int Lv, int flags, const CgEpilogue *cg, sgp_stream_t stream)
{
    SGP_RANGE("sgp_mvm_rows_groups_ex");
    if (!slice_view) return fail(SGP_EINVAL, "sgp_mvm_rows_groups_ex: null view");
    if (Lv < L || (Lv != L && Lv % 4 != 0)) return fail(SGP_EINVAL, "sgp_mvm_rows_groups_ex: Lv must be L or a multiple of 4 above it");
    cudaStream_t st = (cudaStream_t)stream;
    // SGP_MVM_SRC_PADDED (4): src is a zero-padded copy with ceil4(L) columns (lds >= that) of the caller's L-column block, so the
    // splat gathers 16-byte vectors instead of reading ragged rows channel by channel; out keeps its L columns
    const int L_src = (flags & 4) ? (L + 3) / 4 * 4 : L;
    if ((flags & 4) && (lds < L_src || Lv < L_src))
        return fail(SGP_EINVAL, "sgp_mvm_rows_groups_ex: SGP_MVM_SRC_PADDED needs lds >= L rounded up to a multiple of 4 and Lv >= that");
    int rc = (flags & 1) ? sgp_splat_rows_prezeroed(ent, seg_row, n_entries, slice_view->N, slice_view->M, src, lds, L_src, buf0, Lv, stream)
                         : sgp_splat_rows(ent, seg_row, n_entries, slice_view->N, slice_view->M, src, lds, L_src, buf0, Lv, stream);
    if (rc) return rc;
    int in1 = 0;
    rc = sgp_blur_groups(groups, n_groups, slice_view->M, slice_view->order, coeffs, k, Lv, buf0, buf1, &in1,
                         slice_view->fast, stream);
    if (rc) return rc;
    const size_t zero_bytes = sizeof(float) * (size_t)slice_view->M * (size_t)Lv;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaStream_t side = nullptr;
    if ((flags & 2) && in1 && slice_view->M > 0) {   // buf0 is dead: zero it next to the slice
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        side = sgp_side_stream(dev);
        if (side && (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
                     cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess)) {
            if (fork) cudaEventDestroy(fork);
            fork = join = nullptr;
            side = nullptr;
        }
        if (side) {
            CUDA_TRY(cudaEventRecord(fork, st));
            CUDA_TRY(cudaStreamWaitEvent(side, fork, 0));
            CUDA_TRY(cudaMemsetAsync(buf0, 0, zero_bytes, side));
            CUDA_TRY(cudaEventRecord(join, side));
        }
    }
    const float *res = in1 ? buf1 : buf0;
    if (cg && sgp_slice_ring_cg_supported(slice_view, res, Lv, out, ldo, L, src, lds)) {
        rc = sgp_slice_ring_cg(slice_view, res, Lv, out, ldo, L, src, lds, cg->s, cg->noise, cg->pAp, cg->scratch, stream);
    } else {
        rc = sgp_slice(slice_view, res, Lv, out, ldo, L, stream);
        if (cg && !rc) rc = sgp_cg_apply(out, src, cg->s, cg->noise, slice_view->N, L, cg->pAp, cg->scratch, stream);
    }
    if (side) {
        cudaError_t e = cudaStreamWaitEvent(st, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
        if (e != cudaSuccess && !rc) return fail(SGP_ECUDA, "sgp_mvm_rows_groups_ex: %s", cudaGetErrorString(e));
    } else if ((flags & 2) && !rc && slice_view->M > 0) {
        CUDA_TRY(cudaMemsetAsync(buf0, 0, zero_bytes, st));
    }
    return rc;
}

extern "C" int sgp_mvm_rows_groups(const sgp_lattice_view *slice_view, const int32_t *ent, const int32_t *seg_row,
                                   int64_t n_entries, const sgp_blur_group *groups, int n_groups, const float *src, int64_t lds, int L,
                                   const float *coeffs, int k, float *out, int64_t ldo, float *buf0, float *buf1, int Lv,
                                   sgp_stream_t stream)
{
    SGP_RANGE("sgp_mvm_rows_groups");
    if (!slice_view) return fail(SGP_EINVAL, "sgp_mvm_rows_groups: null view");
    if (Lv < L || (Lv != L && Lv % 4 != 0)) return fail(SGP_EINVAL, "sgp_mvm_rows_groups: Lv must be L or a multiple of 4 above it");
    int rc = sgp_splat_rows(ent, seg_row, n_entries, slice_view->N, slice_view->M, src, lds, L, buf0, Lv, stream);
    if (rc) return rc;
    int in1 = 0;
    rc = sgp_blur_groups(groups, n_groups, slice_view->M, slice_view->order, coeffs, k, Lv, buf0, buf1, &in1,
                         slice_view->fast, stream);
    if (rc) return rc;
    return sgp_slice(slice_view, in1 ? buf1 : buf0, Lv, out, ldo, L, stream);
}
