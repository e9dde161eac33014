// sgp_filter.cu -- the reference's native operator in ONE call across the C ABI (B200, sm_100a).
//
//     filter(src[N,L], ref[N,d], coeffs[2r+1]) -> out[N,L]
//
// is what gpytorch_lattice_kernel/cpp/lattice.cpp:6-16 (CPU tensors) and cuda/permutohedral_cuda.cpp:12-22 (CUDA
// tensors) export and bilateral_kernel.py:95,111,119 calls: the lattice of `ref` is built, `src` is splatted, blurred
// along the d+1 axes and sliced, and everything is thrown away again (permutohedral.h:259-340).  sgp_filter is that
// call on device pointers, sgp_filter_host the same on host pointers (the copies ride inside); both run the stage entry
// points of this library in the order a host would: sgp_build_points -> sgp_hash_insert -> sgp_count_points (the one
// host synchronisation: M sizes everything after it) -> sgp_number_points -> sgp_build_neighbours -> sgp_splat ->
// sgp_blur -> sgp_slice.  One product per lattice, so the tables that only pay off over many products (blur groups,
// row-sorted entries) are not built: atomic scatter splat, one blur launch per axis, TMA-ring slice.
//
// The library still allocates no device memory: the caller passes one workspace sized by sgp_filter_workspace_bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail

// one side stream per device (the host variant's second upload; the off-critical-path memset of sgp_mvm_rows_groups_ex):
// created on first use, never destroyed
cudaStream_t sgp_side_stream(int dev)
{
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    if (dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) streams[dev] = nullptr;
    return streams[dev];
}

namespace {

struct FilterWs {
    // phase 1 only (dead once the neighbour table exists) -- the blur buffers reuse this space
    size_t greedy, rank, slot_of, number, table;
    // alive until the product is done
    size_t replay, flags, keys, nbr;
    size_t buf0, buf1;
    // host variant: staged operands
    size_t ref, src, out;
    size_t bytes;
    int64_t cap, M_max;
    size_t number_bytes;
    int Lv;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int filter_layout(int64_t N, int d, int L, int order, int64_t M_max, bool host, FilterWs *w)
{
    if (N < 0 || d < 1 || d > SGP_MAX_DIM || L < 1 || order < 0 || order > SGP_MAX_ORDER)
        return fail(SGP_EINVAL, "sgp_filter: bad shape (N=%lld d=%d L=%d order=%d)", (long long)N, d, L, order);
    const int64_t total = N * (int64_t)(d + 1);
    if ((double)N * (double)(d + 1) >= 4294967295.0) return fail(SGP_EOVERFLOW, "sgp_filter: N*(d+1) does not fit 32 bits");
    if (M_max <= 0 || M_max > total) M_max = total;
    w->M_max = M_max;
    w->cap = sgp_hash_capacity(total);
    w->number_bytes = sgp_number_workspace_bytes(N, d);
    w->Lv = L;
    size_t o = 0;
    // persistent part
    w->replay = o; o += al256((size_t)total * 8);
    w->flags = o; o += 256;
    w->keys = o; o += al256((size_t)M_max * d * 2);
    w->nbr = o; o += al256((size_t)(d + 1) * M_max * 2 * order * 4);
    if (host) {
        w->ref = o; o += al256((size_t)N * d * 4);
        w->src = o; o += al256((size_t)N * L * 4);
        w->out = o; o += al256((size_t)N * L * 4);
    } else {
        w->ref = w->src = w->out = 0;
    }
    // phase 1 / phase 2 share the rest
    const size_t base = o;
    size_t a = base;
    w->greedy = a; a += al256((size_t)total * 2);
    w->rank = a; a += al256((size_t)total);
    w->slot_of = a; a += al256((size_t)total * 4);
    w->number = a; a += al256(w->number_bytes);
    w->table = a; a += al256((size_t)w->cap * 8);
    size_t b = base;
    w->buf0 = b; b += al256((size_t)M_max * L * 4);
    w->buf1 = b; b += al256((size_t)M_max * L * 4);
    w->bytes = a > b ? a : b;
    return SGP_OK;
}

// [N, width] fp32 rows between pitched buffers: one flat copy when both sides are dense (the common case: the 2-D form
// takes a slower path for large pinned transfers)
cudaError_t copy_rows(float *dst, int64_t ldd, const float *src, int64_t lds, int width, int64_t N, cudaMemcpyKind kind,
                      cudaStream_t st)
{
    if (ldd == width && lds == width) return cudaMemcpyAsync(dst, src, (size_t)N * width * 4, kind, st);
    return cudaMemcpy2DAsync(dst, (size_t)ldd * 4, src, (size_t)lds * 4, (size_t)width * 4, (size_t)N, kind, st);
}

cudaStream_t copy_stream(int dev) { return sgp_side_stream(dev); }

// How the host entry pipelines the product with its copies: the rows of src arrive in `n_chunks` pieces (ready[c] fires
// when piece c is on the device) and are splatted as they come; after the blur the slice runs piece by piece and each
// piece of the result starts its way back (on `down`) while the next one is still being sliced.
// first point of piece c of C: multiples of 32 points, so that a piece's rows of the replay table, of src and of out keep
// the 16-byte alignment the kernels' vector loads count on
inline int64_t filter_chunk_begin(int64_t N, int c, int C)
{
    if (c >= C) return N;
    return ((N * (int64_t)c) / C) & ~(int64_t)31;
}

struct FilterPipe {
    int n_chunks;
    const cudaEvent_t *ready;   // [n_chunks]
    cudaStream_t down;          // stream of the device -> host copies
    float *out_host;
    int64_t ldo_host;
    cudaEvent_t sliced;         // scratch event
};

int filter_device(const float *src, int64_t lds, const float *ref, int64_t ldx, const float *coeffs, int k, int64_t N,
                  int L, int d, float *out, int64_t ldo, char *base, const FilterWs &w, int64_t *M_out, cudaStream_t st,
                  const FilterPipe *pipe)
{
    const int order = k / 2;
    float var = 0.0f;
    int rc = sgp_stencil_variance(coeffs, k, &var);
    if (rc) return rc;
    float scale[SGP_MAX_DIM];
    rc = sgp_scale_factors(d, var, scale);
    if (rc) return rc;
    int16_t *greedy = (int16_t *)(base + w.greedy);
    int8_t *rank = (int8_t *)(base + w.rank);
    int32_t *replay = (int32_t *)(base + w.replay);
    int32_t *flags = (int32_t *)(base + w.flags);
    uint64_t *table = (uint64_t *)(base + w.table);
    uint32_t *slot_of = (uint32_t *)(base + w.slot_of);
    int16_t *keys = (int16_t *)(base + w.keys);
    int32_t *nbr = (int32_t *)(base + w.nbr);
    float *buf0 = (float *)(base + w.buf0), *buf1 = (float *)(base + w.buf1);
    sgp_stream_t s = (sgp_stream_t)st;
    CUDA_TRY(cudaMemsetAsync(flags, 0, 16, st));
    CUDA_TRY(cudaMemsetAsync(table, 0xFF, (size_t)w.cap * 8, st));
    rc = sgp_build_points(ref, N, d, ldx, scale, greedy, rank, replay, flags, s);
    if (rc) return rc;
    rc = sgp_hash_insert(greedy, rank, N, d, table, w.cap, slot_of, flags, s);
    if (rc) return rc;
    int64_t M = 0;
    int32_t fl = 0;
    rc = sgp_count_points(table, w.cap, slot_of, N, d, base + w.number, w.number_bytes, flags, &M, &fl, s);
    if (M_out) *M_out = M;
    if (rc) return rc;
    if (M > w.M_max)
        return fail(SGP_ENOMEM, "sgp_filter: the lattice has %lld points but the workspace was sized for %lld", (long long)M,
                    (long long)w.M_max);
    rc = sgp_number_points(table, w.cap, slot_of, greedy, rank, N, d, base + w.number, M, replay, keys, s);
    if (rc) return rc;
    rc = sgp_build_neighbours(keys, M, d, order, table, w.cap, nbr, s);
    if (rc) return rc;
    sgp_lattice_view v;
    memset(&v, 0, sizeof(v));
    v.N = N;
    v.M = M;
    v.d = d;
    v.order = order;
    v.replay = replay;
    v.nbr = order > 0 ? nbr : nullptr;
    v.fast = 1;
    if (!pipe) {
        rc = sgp_splat(&v, src, lds, L, buf0, SGP_SPLAT_ATOMIC, s);   // (buf0/buf1 overwrite the build scratch: stream order)
        if (rc) return rc;
        int in1 = 0;
        rc = sgp_blur(&v, coeffs, k, L, buf0, buf1, &in1, s);
        if (rc) return rc;
        return sgp_slice(&v, in1 ? buf1 : buf0, L, out, ldo, L, s);
    }
    // pipelined with the copies: views of point ranges (the replay table is [N, d+1, 2], point-major)
    const int C = pipe->n_chunks;
    auto chunk_begin = [&](int c) { return filter_chunk_begin(N, c, C); };
    CUDA_TRY(cudaMemsetAsync(buf0, 0, sizeof(float) * (size_t)M * (size_t)L, st));
    for (int c = 0; c < C; ++c) {
        const int64_t p0 = chunk_begin(c), p1 = chunk_begin(c + 1);
        if (p1 == p0) continue;
        CUDA_TRY(cudaStreamWaitEvent(st, pipe->ready[c], 0));
        sgp_lattice_view vc = v;
        vc.N = p1 - p0;
        vc.replay = replay + p0 * (int64_t)(d + 1) * 2;
        rc = sgp_splat(&vc, src + p0 * lds, lds, L, buf0, SGP_SPLAT_ATOMIC_ACCUMULATE, s);
        if (rc) return rc;
    }
    int in1 = 0;
    rc = sgp_blur(&v, coeffs, k, L, buf0, buf1, &in1, s);
    if (rc) return rc;
    for (int c = 0; c < C; ++c) {
        const int64_t p0 = chunk_begin(c), p1 = chunk_begin(c + 1);
        if (p1 == p0) continue;
        sgp_lattice_view vc = v;
        vc.N = p1 - p0;
        vc.replay = replay + p0 * (int64_t)(d + 1) * 2;
        rc = sgp_slice(&vc, in1 ? buf1 : buf0, L, out + p0 * ldo, ldo, L, s);
        if (rc) return rc;
        // this piece of the result leaves while the next one is sliced (stream order on `down` keeps the event reusable)
        CUDA_TRY(cudaEventRecord(pipe->sliced, st));
        CUDA_TRY(cudaStreamWaitEvent(pipe->down, pipe->sliced, 0));
        CUDA_TRY(copy_rows(pipe->out_host + p0 * pipe->ldo_host, pipe->ldo_host, out + p0 * ldo, ldo, L, p1 - p0,
                           cudaMemcpyDeviceToHost, pipe->down));
    }
    return SGP_OK;
}

}   // namespace

extern "C" size_t sgp_filter_workspace_bytes(int64_t N, int d, int L, int order, int64_t M_max)
{
    FilterWs w;
    if (filter_layout(N, d, L, order, M_max, false, &w) != SGP_OK) return 0;
    return w.bytes;
}

extern "C" size_t sgp_filter_host_workspace_bytes(int64_t N, int d, int L, int order, int64_t M_max)
{
    FilterWs w;
    if (filter_layout(N, d, L, order, M_max, true, &w) != SGP_OK) return 0;
    return w.bytes;
}

extern "C" int sgp_filter(const float *src, int64_t lds, const float *ref, int64_t ldx, const float *coeffs, int k,
                          int64_t N, int L, int d, float *out, int64_t ldo, void *workspace, size_t workspace_bytes,
                          int64_t M_max, int64_t *M_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_filter");
    if (M_out) *M_out = 0;
    if (!coeffs || k < 1 || (k & 1) == 0) return fail(SGP_EINVAL, "sgp_filter: the stencil must have odd length >= 1");
    FilterWs w;
    int rc = filter_layout(N, d, L, k / 2, M_max, false, &w);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!src || !ref || !out || !workspace || lds < L || ldo < L || ldx < d)
        return fail(SGP_EINVAL, "sgp_filter: null pointer or a leading dimension below its width");
    if (workspace_bytes < w.bytes)
        return fail(SGP_EINVAL, "sgp_filter: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    if (((uintptr_t)workspace & 255) != 0) return fail(SGP_EINVAL, "sgp_filter: the workspace must be 256-byte aligned");
    return filter_device(src, lds, ref, ldx, coeffs, k, N, L, d, out, ldo, (char *)workspace, w, M_out,
                         (cudaStream_t)stream, nullptr);
}

extern "C" int sgp_filter_host(const float *src_host, int64_t lds, const float *ref_host, int64_t ldx, const float *coeffs,
                               int k, int64_t N, int L, int d, float *out_host, int64_t ldo, void *workspace,
                               size_t workspace_bytes, int64_t M_max, int64_t *M_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_filter_host");
    if (M_out) *M_out = 0;
    if (!coeffs || k < 1 || (k & 1) == 0) return fail(SGP_EINVAL, "sgp_filter_host: the stencil must have odd length >= 1");
    FilterWs w;
    int rc = filter_layout(N, d, L, k / 2, M_max, true, &w);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!src_host || !ref_host || !out_host || !workspace || lds < L || ldo < L || ldx < d)
        return fail(SGP_EINVAL, "sgp_filter_host: null pointer or a leading dimension below its width");
    if (workspace_bytes < w.bytes)
        return fail(SGP_EINVAL, "sgp_filter_host: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    if (((uintptr_t)workspace & 255) != 0) return fail(SGP_EINVAL, "sgp_filter_host: the workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)workspace;
    float *ref_d = (float *)(base + w.ref), *src_d = (float *)(base + w.src), *out_d = (float *)(base + w.out);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    // positions first (the lattice build needs them); the RHS block travels on a second stream while the lattice is
    // being built.  Pinned host memory is copied asynchronously, pageable memory through the driver's staging.
    CUDA_TRY(copy_rows(ref_d, d, ref_host, ldx, d, N, cudaMemcpyHostToDevice, st));
    // ... in SGP_FILTER_CHUNKS pieces of rows (default 4), each splatted as soon as it has arrived; after the blur the
    // result is sliced piece by piece and every piece starts its way back while the next one is sliced: the product
    // hides behind the copies except for the blur.
    cudaStream_t side = copy_stream(dev);
    constexpr int MAX_CHUNKS = 16;
    int n_chunks = 4;
    if (const char *e = getenv("SGP_FILTER_CHUNKS")) n_chunks = atoi(e);
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > MAX_CHUNKS) n_chunks = MAX_CHUNKS;
    if ((int64_t)n_chunks > N) n_chunks = (int)N;
    cudaEvent_t fork = nullptr, sliced = nullptr, ready[MAX_CHUNKS] = {};
    bool ok = side != nullptr && cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&sliced, cudaEventDisableTiming) == cudaSuccess;
    for (int c = 0; ok && c < n_chunks; ++c) ok = cudaEventCreateWithFlags(&ready[c], cudaEventDisableTiming) == cudaSuccess;
    auto destroy_events = [&]() {
        if (fork) cudaEventDestroy(fork);
        if (sliced) cudaEventDestroy(sliced);
        for (int c = 0; c < MAX_CHUNKS; ++c)
            if (ready[c]) cudaEventDestroy(ready[c]);
    };
    cudaError_t ce = cudaSuccess;
    if (!ok) {
        // no second stream / events: everything in order on the caller's stream
        destroy_events();
        ce = copy_rows(src_d, L, src_host, lds, L, N, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) rc = filter_device(src_d, L, ref_d, d, coeffs, k, N, L, d, out_d, L, base, w, M_out, st, nullptr);
        if (ce == cudaSuccess && rc == SGP_OK) ce = copy_rows(out_host, ldo, out_d, L, L, N, cudaMemcpyDeviceToHost, st);
        cudaError_t se0 = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) return fail(SGP_ECUDA, "sgp_filter_host: copy failed: %s", cudaGetErrorString(ce));
        if (rc) return rc;
        if (se0 != cudaSuccess) return fail(SGP_ECUDA, "sgp_filter_host: %s", cudaGetErrorString(se0));
        return SGP_OK;
    }
    // the workspace may still be in use by earlier work of the caller's stream
    ce = cudaEventRecord(fork, st);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(side, fork, 0);
    for (int c = 0; ce == cudaSuccess && c < n_chunks; ++c) {
        const int64_t p0 = filter_chunk_begin(N, c, n_chunks), p1 = filter_chunk_begin(N, c + 1, n_chunks);
        if (p1 > p0)
            ce = copy_rows(src_d + p0 * L, L, src_host + p0 * lds, lds, L, p1 - p0, cudaMemcpyHostToDevice, side);
        if (ce == cudaSuccess) ce = cudaEventRecord(ready[c], side);
    }
    if (ce == cudaSuccess) {
        FilterPipe pipe{n_chunks, ready, side, out_host, ldo, sliced};
        rc = filter_device(src_d, L, ref_d, d, coeffs, k, N, L, d, out_d, L, base, w, M_out, st, &pipe);
    }
    cudaStreamSynchronize(side);   // also on the error paths: neither the uploads nor the downloads may outlive the call
    cudaError_t se = cudaStreamSynchronize(st);
    if (se == cudaSuccess) se = cudaStreamSynchronize(side);   // (downloads enqueued behind work of st)
    destroy_events();
    if (ce != cudaSuccess) return fail(SGP_ECUDA, "sgp_filter_host: copy failed: %s", cudaGetErrorString(ce));
    if (rc) return rc;
    if (se != cudaSuccess) return fail(SGP_ECUDA, "sgp_filter_host: %s", cudaGetErrorString(se));
    return SGP_OK;
}
