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
//   lnb[p][a][t]    position, relative to batch_begin[b], of the neighbour along the group's a-th axis, offset
//                   t over o = -r..-1, 1..r; 0xFFFF = absent
// The arithmetic of a pass is that of sgp_blur_kernel (same products, same order, no FMA), so results are
// bit-identical to the per-axis path.
//
// Sorting / scanning in the build uses CUB (CUDA toolkit header library); the MVM kernel is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok
#define grid_for sgp_grid_for

#define GROUP_THREADS 512
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
sgp_group_hash_kernel(const int16_t *__restrict__ keys, int64_t M, int d, int j0, int j1,
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
    hash[i] = h;
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

// batch b owns the classes that START in [b*window, (b+1)*window): batch_begin[b] = first p with class_start[p] >= b*window
__global__ void __launch_bounds__(256)
sgp_group_batches_kernel(const uint32_t *__restrict__ class_start, int64_t M, int64_t window, int64_t n_batches,
                         uint32_t *__restrict__ batch_begin, uint32_t *__restrict__ max_rows)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_batches) return;
    auto lower = [&](int64_t target) {
        int64_t lo = 0, hi = M;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)class_start[mid] >= target) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    const int64_t begin = (b == n_batches) ? M : lower(b * window);
    batch_begin[b] = (uint32_t)begin;
    if (b < n_batches) {
        const int64_t end = (b + 1 == n_batches) ? M : lower((b + 1) * window);
        atomicMax(max_rows, (uint32_t)(end - begin));
    }
}

__global__ void __launch_bounds__(256)
sgp_group_tables_kernel(const int32_t *__restrict__ nbr, int64_t M, int order_r, int j0, int j1,
                        const uint32_t *__restrict__ order, const uint32_t *__restrict__ pos,
                        const uint32_t *__restrict__ class_start, const uint32_t *__restrict__ prev_pos, int64_t window,
                        const uint32_t *__restrict__ batch_begin, int32_t *__restrict__ src, uint16_t *__restrict__ lnb,
                        int32_t *__restrict__ flags)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const uint32_t row = order[p];
    src[p] = prev_pos ? (int32_t)prev_pos[row] : (int32_t)row;
    const uint32_t base = batch_begin[class_start[p] / window];
    const int nax = j1 - j0, w = 2 * order_r;
    for (int a = 0; a < nax; ++a) {
        const int32_t *np = nbr + ((int64_t)(j0 + a) * M + row) * w;
        for (int t = 0; t < w; ++t) {
            const int32_t nb = np[t];
            uint32_t local = LNB_ABSENT;
            if (nb >= 0) {
                const uint32_t q = pos[nb];
                local = q - base;
                if (q < base || local >= LNB_ABSENT) {   // neighbour outside the CTA's rows: must not happen
                    atomicOr(flags, 4);
                    local = LNB_ABSENT;
                }
            }
            lnb[(p * nax + a) * w + t] = (uint16_t)local;
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

extern "C" size_t sgp_group_workspace_bytes(int64_t M)
{
    GroupWs w;
    if (M <= 0 || group_ws_layout(M, &w) != SGP_OK) return 0;
    return w.bytes;
}

extern "C" int sgp_group_prepare(const int16_t *keys, int64_t M, int d, int j0, int j1, uint32_t *order, uint32_t *pos,
                                 uint32_t *class_start, void *workspace, size_t workspace_bytes, int64_t *max_class_out,
                                 sgp_stream_t stream)
{
    if (!keys || !order || !pos || !class_start || !workspace || !max_class_out || M <= 0 || d < 1 || d > SGP_MAX_DIM ||
        j0 < 0 || j1 <= j0 || j1 > d + 1)
        return fail(SGP_EINVAL, "sgp_group_prepare: bad argument");
    if (M >= (1ll << 32)) return fail(SGP_EOVERFLOW, "M does not fit 32 bits");
    GroupWs w;
    int rc = group_ws_layout(M, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "group workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)workspace;
    unsigned long long *ha = (unsigned long long *)(base + w.hash_a), *hb = (unsigned long long *)(base + w.hash_b);
    uint32_t *ids = (uint32_t *)(base + w.ids_a), *head = (uint32_t *)(base + w.head);
    uint32_t *small = (uint32_t *)(base + w.small);
    size_t cub_bytes = w.cub_bytes;
    sgp_group_hash_kernel<<<grid_for(M, 256), 256, 0, st>>>(keys, M, d, j0, j1, ha, ids);
    rc = launch_ok("sgp_group_hash_kernel");
    if (rc) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(base + w.cub, cub_bytes, ha, hb, ids, order, (int64_t)M, 0, 64, st));
    sgp_group_heads_kernel<<<grid_for(M, 256), 256, 0, st>>>(hb, M, head);
    rc = launch_ok("sgp_group_heads_kernel");
    if (rc) return rc;
    cub_bytes = w.cub_bytes;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(base + w.cub, cub_bytes, head, class_start, cub::Max(), (int64_t)M, st));
    CUDA_TRY(cudaMemsetAsync(small, 0, 16, st));
    sgp_group_pos_kernel<<<grid_for(M, 256), 256, 0, st>>>(order, class_start, M, pos, small);
    rc = launch_ok("sgp_group_pos_kernel");
    if (rc) return rc;
    uint32_t mx = 0;
    CUDA_TRY(cudaMemcpyAsync(&mx, small, sizeof(mx), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *max_class_out = (int64_t)mx;
    return SGP_OK;
}

extern "C" int sgp_group_finalize(const int32_t *nbr, int64_t M, int order_r, int j0, int j1, const uint32_t *order,
                                  const uint32_t *pos, const uint32_t *class_start, const uint32_t *prev_pos,
                                  int64_t window, int64_t n_batches, uint32_t *batch_begin, int32_t *src, uint16_t *lnb,
                                  void *workspace, size_t workspace_bytes, int32_t *max_rows_out, sgp_stream_t stream)
{
    if (!nbr || !order || !pos || !class_start || !batch_begin || !src || !lnb || !workspace || !max_rows_out || M <= 0 ||
        order_r < 1 || order_r > SGP_MAX_ORDER || j1 <= j0 || window < 1 || n_batches != (M + window - 1) / window)
        return fail(SGP_EINVAL, "sgp_group_finalize: bad argument");
    GroupWs w;
    int rc = group_ws_layout(M, &w);
    if (rc) return rc;
    if (workspace_bytes < w.bytes) return fail(SGP_EINVAL, "group workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *small = (uint32_t *)((char *)workspace + w.small);
    CUDA_TRY(cudaMemsetAsync(small, 0, 16, st));
    sgp_group_batches_kernel<<<grid_for(n_batches + 1, 256), 256, 0, st>>>(class_start, M, window, n_batches, batch_begin,
                                                                            small);
    rc = launch_ok("sgp_group_batches_kernel");
    if (rc) return rc;
    sgp_group_tables_kernel<<<grid_for(M, 256), 256, 0, st>>>(nbr, M, order_r, j0, j1, order, pos, class_start, prev_pos,
                                                               window, batch_begin, src, lnb, (int32_t *)(small + 1));
    rc = launch_ok("sgp_group_tables_kernel");
    if (rc) return rc;
    uint32_t host[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(host, small, sizeof(host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (host[1] != 0) return fail(SGP_EINVAL, "blur group [%d,%d): a neighbour fell outside its CTA batch", j0, j1);
    *max_rows_out = (int32_t)host[0];
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
// the MVM kernel: one CTA per batch (x one block of CB channels)
// ------------------------------------------------------------------------------------
struct GroupCoeffs {
    float c[2 * SGP_MAX_ORDER + 1];
};

#define STAGE_UNROLL 4

template <int VEC, int R>
__global__ void __launch_bounds__(GROUP_THREADS)
sgp_blur_group_kernel(const uint32_t *__restrict__ batch_begin, const int32_t *__restrict__ src,
                      const uint16_t *__restrict__ lnb, const float *__restrict__ in, float *__restrict__ out, int L,
                      int CB, int rows_cap, int nax, int order_rt, GroupCoeffs cf)
{
    constexpr int RR = R > 0 ? R : SGP_MAX_ORDER;
    const int r = R > 0 ? R : order_rt;
    extern __shared__ __align__(16) float smem[];
    float *A = smem, *B = smem + (size_t)rows_cap * CB;
    const uint32_t p0 = batch_begin[blockIdx.x];
    const int rows = (int)(batch_begin[blockIdx.x + 1] - p0);
    if (rows == 0) return;
    const int cb0 = blockIdx.y * CB;
    const int cb = min(CB, L - cb0);
    const int chunks = cb / VEC;
    const int items = rows * chunks;

    // stage: gather the batch's rows from the previous stage's order; STAGE_UNROLL independent row loads in flight
    for (int w0 = threadIdx.x; w0 < items; w0 += GROUP_THREADS * STAGE_UNROLL) {
        int srow[STAGE_UNROLL];
        Vec<VEC> v[STAGE_UNROLL];
#pragma unroll
        for (int u = 0; u < STAGE_UNROLL; ++u) {
            const int w = w0 + u * GROUP_THREADS;
            srow[u] = (w < items) ? __ldg(src + p0 + w / chunks) : -1;
        }
#pragma unroll
        for (int u = 0; u < STAGE_UNROLL; ++u) {
            const int w = w0 + u * GROUP_THREADS;
            if (srow[u] >= 0) v[u].load_cg(in + (int64_t)srow[u] * L + cb0 + (w % chunks) * VEC);
        }
#pragma unroll
        for (int u = 0; u < STAGE_UNROLL; ++u) {
            const int w = w0 + u * GROUP_THREADS;
            if (srow[u] >= 0) v[u].store(A + (w / chunks) * CB + (w % chunks) * VEC);
        }
    }
    __syncthreads();

    const int w2 = 2 * r;
    for (int a = 0; a < nax; ++a) {
        for (int w = threadIdx.x; w < items; w += GROUP_THREADS) {
            const int lr = w / chunks, c = (w - lr * chunks) * VEC;
            const uint16_t *nb = lnb + ((int64_t)(p0 + lr) * nax + a) * w2;
            uint32_t ni[2 * RR];
#pragma unroll
            for (int t = 0; t < 2 * RR; ++t) ni[t] = (t < w2) ? (uint32_t)__ldg(nb + t) : LNB_ABSENT;
            Vec<VEC> acc;
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
            // reference order: o = -r..-1, 0, 1..r
#pragma unroll
            for (int t = 0; t < RR; ++t) {
                if (t < r && ni[t] != LNB_ABSENT) {
                    Vec<VEC> v;
                    v.load_plain(A + ni[t] * CB + c);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc.v[k] = __fadd_rn(acc.v[k], __fmul_rn(cf.c[t], v.v[k]));
                }
            }
            {
                Vec<VEC> v;
                v.load_plain(A + lr * CB + c);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc.v[k] = __fadd_rn(acc.v[k], __fmul_rn(cf.c[r], v.v[k]));
            }
#pragma unroll
            for (int t = 0; t < RR; ++t) {
                if (t < r && ni[r + t] != LNB_ABSENT) {
                    Vec<VEC> v;
                    v.load_plain(A + ni[r + t] * CB + c);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc.v[k] = __fadd_rn(acc.v[k], __fmul_rn(cf.c[r + 1 + t], v.v[k]));
                }
            }
            acc.store(B + lr * CB + c);
        }
        __syncthreads();
        float *t = A; A = B; B = t;
    }

    for (int w = threadIdx.x; w < items; w += GROUP_THREADS) {
        const int lr = w / chunks, c = (w - lr * chunks) * VEC;
        Vec<VEC> v;
        v.load_plain(A + lr * CB + c);
        v.store(out + (int64_t)(p0 + lr) * L + cb0 + c);
    }
}

template <int VEC>
static int launch_group(const sgp_blur_group *g, int order, const GroupCoeffs &cf, int L, int CB, const float *in,
                        float *out, cudaStream_t st)
{
    const size_t smem = (size_t)2 * g->rows_cap * CB * sizeof(float);
    dim3 grid((unsigned)g->n_batches, (unsigned)((L + CB - 1) / CB));
    const int nax = g->j1 - g->j0;
#define SGP_LAUNCH_GROUP(RR)                                                                                          \
    do {                                                                                                              \
        if (smem > 48 * 1024)                                                                                         \
            CUDA_TRY(cudaFuncSetAttribute(sgp_blur_group_kernel<VEC, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                                \
        sgp_blur_group_kernel<VEC, RR><<<grid, GROUP_THREADS, smem, st>>>(g->batch_begin, g->src, g->lnb, in, out, L,   \
                                                                         CB, g->rows_cap, nax, order, cf);            \
    } while (0)
    if (order == 1) SGP_LAUNCH_GROUP(1);
    else if (order == 2) SGP_LAUNCH_GROUP(2);
    else if (order == 3) SGP_LAUNCH_GROUP(3);
    else SGP_LAUNCH_GROUP(0);
#undef SGP_LAUNCH_GROUP
    return launch_ok("sgp_blur_group_kernel");
}

extern "C" int sgp_blur_groups_channel_block(int L) { return L < 16 ? L : 16; }

extern "C" int sgp_blur_groups(const sgp_blur_group *groups, int n_groups, int64_t M, int order, const float *coeffs,
                               int k, int L, float *buf0, float *buf1, int *result_in_buf1, sgp_stream_t stream)
{
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
    if (L % 4 == 0 && CB % 4 == 0 && al(buf0, 16) && al(buf1, 16)) vec = 4;
    else if (L % 2 == 0 && CB % 2 == 0 && al(buf0, 8) && al(buf1, 8)) vec = 2;
    float *in = buf0, *out = buf1;
    for (int gi = 0; gi < n_groups; ++gi) {
        const sgp_blur_group *g = groups + gi;
        if (!g->batch_begin || !g->src || !g->lnb || g->rows_cap < 1 || g->n_batches < 1 || g->j1 <= g->j0)
            return fail(SGP_EINVAL, "sgp_blur_groups: group %d is not built", gi);
        if ((size_t)2 * g->rows_cap * CB * sizeof(float) > 227 * 1024)
            return fail(SGP_EUNSUPPORTED, "blur group %d needs %d rows x %d channels of shared memory", gi, g->rows_cap, CB);
        int rc;
        if (vec == 4) rc = launch_group<4>(g, order, cf, L, CB, in, out, st);
        else if (vec == 2) rc = launch_group<2>(g, order, cf, L, CB, in, out, st);
        else rc = launch_group<1>(g, order, cf, L, CB, in, out, st);
        if (rc) return rc;
        float *t = in; in = out; out = t;
    }
    if (result_in_buf1) *result_in_buf1 = (in == buf1) ? 1 : 0;
    return SGP_OK;
}
