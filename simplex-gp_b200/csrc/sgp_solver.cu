// sgp_solver.cu -- the vector updates of batched conjugate gradients, either side of the lattice MVM (B200, sm_100a).
//
// The reference hands the solve of (s K + noise I) X = B to GPyTorch (experiments/train_simplexgp.py:29-48: CG on
// [y - mu | probes], Lanczos coefficients for the log-determinant).  One CG iteration is one lattice MVM plus three
// sweeps over [N, L] blocks; written with tensor expressions those sweeps are ~17 small launches that cost more device
// time than the MVM itself at N = 1M, L = 11.  Here each sweep is one launch with its column-wise dot products folded in:
//
//   sgp_cg_apply      AP <- s * KP + noise * P  (in place on the MVM's output)      pAp[l] = sum_n P * AP
//   sgp_cg_update     alpha = rs / pAp;  X += alpha P;  R -= alpha AP               rs_new[l] = sum_n R * R
//   sgp_cg_direction  beta = rs_new / rs;  P <- R + beta P;  rs <- rs_new           done = all(sqrt(rs_new) / bnorm < tol)
//
// Blocks are [N, L] row-major with unit column stride and no row padding (ld = L).  A thread walks the flattened block
// with a stride that is a multiple of L, so it always sees the same column and keeps its partial dot product in a
// register; partials go to a [blocks, L] scratch and a one-block second stage sums them in a fixed order: the results
// do not depend on scheduling.  s and noise are read from device memory (they are parameters of the model; no host
// round trip).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok

#define CG_THREADS 256
#define CG_MAX_BLOCKS 1184         /* 8 per SM */
#define CG_UNROLL 4                /* independent elements per thread and trip: enough loads in flight for HBM */
#define CG_MAX_COLUMNS CG_THREADS

struct CgGeometry {
    int active;      // threads that take part: the largest multiple of L not above CG_THREADS
    int blocks;
    int64_t per_block;   // elements per block, a multiple of `active`
};

static CgGeometry cg_geometry(int64_t N, int L)
{
    CgGeometry g;
    g.active = (CG_THREADS / L) * L;
    const int64_t total = N * (int64_t)L;
    int64_t blocks = (total + (int64_t)g.active * 2 * CG_UNROLL - 1) / ((int64_t)g.active * 2 * CG_UNROLL);
    if (blocks > CG_MAX_BLOCKS) blocks = CG_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    int64_t per = (total + blocks - 1) / blocks;
    per = (per + g.active - 1) / g.active * g.active;
    g.blocks = (int)((total + per - 1) / per);
    g.per_block = per;
    return g;
}

// sum the per-thread partials of equal column (thread t holds column t % L) and store them at partial[block, :]
__device__ __forceinline__ void cg_block_columns(float acc, int active, int L, float *__restrict__ partial)
{
    __shared__ float s_acc[CG_THREADS];
    s_acc[threadIdx.x] = threadIdx.x < active ? acc : 0.0f;
    __syncthreads();
    if (threadIdx.x < L) {
        float t = 0.0f;
        for (int k = threadIdx.x; k < active; k += L) t += s_acc[k];
        partial[(int64_t)blockIdx.x * L + threadIdx.x] = t;
    }
}

// second stage helper (one block): thread l < L returns sum_b partial[b, l].  All `active` threads take part -- thread t
// sums the blocks b = t / L, t / L + active / L, ... of column t % L -- and the per-thread sums are combined in a fixed
// order through shared memory.
__device__ __forceinline__ float cg_sum_partials(const float *__restrict__ partial, int blocks, int L, int active)
{
    __shared__ float s_part[CG_THREADS];
    float t = 0.0f;
    if (threadIdx.x < active) {
        const int col = threadIdx.x % L, step = active / L;
        // eight loads in flight per thread (a dependent chain of ~50 L2 round trips took 13 us); the eight partial sums
        // are combined in a fixed order
        float u8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        int b = threadIdx.x / L;
        for (; b + 7 * step < blocks; b += 8 * step) {
            float v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v8[u] = partial[(int64_t)(b + u * step) * L + col];
#pragma unroll
            for (int u = 0; u < 8; ++u) u8[u] += v8[u];
        }
        for (; b < blocks; b += step) u8[0] += partial[(int64_t)b * L + col];
        t = ((u8[0] + u8[1]) + (u8[2] + u8[3])) + ((u8[4] + u8[5]) + (u8[6] + u8[7]));
    }
    s_part[threadIdx.x] = t;
    __syncthreads();
    float total = 0.0f;
    if (threadIdx.x < L)
        for (int k = threadIdx.x; k < active; k += L) total += s_part[k];
    return total;
}

__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_apply_kernel(float *__restrict__ AP, const float *__restrict__ P, const float *__restrict__ s_ptr,
                    const float *__restrict__ noise_ptr, int64_t total, int L, int active, int64_t per_block,
                    float *__restrict__ partial)
{
    const float s = __ldg(s_ptr), noise = __ldg(noise_ptr);
    float acc = 0.0f;
    if (threadIdx.x < active) {
        const int64_t lo = (int64_t)blockIdx.x * per_block;
        const int64_t hi = min(lo + per_block, total);
        int64_t i = lo + threadIdx.x;
        for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
            float p[CG_UNROLL], kp[CG_UNROLL];
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) { p[u] = __ldcs(P + i + (int64_t)u * active); kp[u] = __ldcs(AP + i + (int64_t)u * active); }
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
                const float ap = fmaf(s, kp[u], noise * p[u]);
                AP[i + (int64_t)u * active] = ap;
                acc = fmaf(p[u], ap, acc);
            }
        }
        for (; i < hi; i += active) {
            const float p = P[i];
            const float ap = fmaf(s, AP[i], noise * p);
            AP[i] = ap;
            acc = fmaf(p, ap, acc);
        }
    }
    cg_block_columns(acc, active, L, partial);
}

__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_reduce_kernel(const float *__restrict__ partial, int blocks, int L, int active, float *__restrict__ out)
{
    const float t = cg_sum_partials(partial, blocks, L, active);
    if (threadIdx.x < L) out[threadIdx.x] = t;
}

__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_update_kernel(float *__restrict__ X, float *__restrict__ R, const float *__restrict__ P,
                     const float *__restrict__ AP, const float *__restrict__ rs, const float *__restrict__ pAp,
                     int64_t total, int L, int active, int64_t per_block, float *__restrict__ alpha_out,
                     float *__restrict__ partial)
{
    float acc = 0.0f;
    if (threadIdx.x < active) {
        const int col = threadIdx.x % L;
        const float alpha = __ldg(rs + col) / fmaxf(__ldg(pAp + col), 1e-30f);
        if (blockIdx.x == 0 && threadIdx.x < L) alpha_out[col] = alpha;
        const int64_t lo = (int64_t)blockIdx.x * per_block;
        const int64_t hi = min(lo + per_block, total);
        int64_t i = lo + threadIdx.x;
        for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
            float xv[CG_UNROLL], rv[CG_UNROLL], pv[CG_UNROLL], av[CG_UNROLL];
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
                const int64_t q = i + (int64_t)u * active;
                xv[u] = __ldcs(X + q); rv[u] = __ldcs(R + q); pv[u] = __ldcs(P + q); av[u] = __ldcs(AP + q);
            }
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
                const int64_t q = i + (int64_t)u * active;
                X[q] = fmaf(alpha, pv[u], xv[u]);
                const float r = fmaf(-alpha, av[u], rv[u]);
                R[q] = r;
                acc = fmaf(r, r, acc);
            }
        }
        for (; i < hi; i += active) {
            X[i] = fmaf(alpha, P[i], X[i]);
            const float r = fmaf(-alpha, AP[i], R[i]);
            R[i] = r;
            acc = fmaf(r, r, acc);
        }
    }
    cg_block_columns(acc, active, L, partial);
}

// one block: rs_new from the partials, beta, the convergence flag; rs <- rs_new
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_beta_kernel(const float *__restrict__ partial, int blocks, int L, int active, float *__restrict__ rs,
                   const float *__restrict__ bnorm, float tol, float *__restrict__ beta_out, int32_t *__restrict__ done)
{
    __shared__ int s_open;
    if (threadIdx.x == 0) s_open = 0;
    const float rs_new = cg_sum_partials(partial, blocks, L, active);   // contains a __syncthreads
    __syncthreads();
    if (threadIdx.x < L) {
        beta_out[threadIdx.x] = rs_new / fmaxf(rs[threadIdx.x], 1e-30f);
        rs[threadIdx.x] = rs_new;
        if (!(sqrtf(rs_new) / bnorm[threadIdx.x] < tol)) atomicAdd(&s_open, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) *done = s_open == 0 ? 1 : 0;
}

__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_direction_kernel(float *__restrict__ P, const float *__restrict__ R, const float *__restrict__ beta,
                        int64_t total, int L, int active, int64_t per_block)
{
    if (threadIdx.x >= active) return;
    const float b = __ldg(beta + threadIdx.x % L);
    const int64_t lo = (int64_t)blockIdx.x * per_block;
    const int64_t hi = min(lo + per_block, total);
    int64_t i = lo + threadIdx.x;
    for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
        float pv[CG_UNROLL], rv[CG_UNROLL];
#pragma unroll
        for (int u = 0; u < CG_UNROLL; ++u) { pv[u] = __ldcs(P + i + (int64_t)u * active); rv[u] = __ldcs(R + i + (int64_t)u * active); }
#pragma unroll
        for (int u = 0; u < CG_UNROLL; ++u) P[i + (int64_t)u * active] = fmaf(b, pv[u], rv[u]);
    }
    for (; i < hi; i += active) P[i] = fmaf(b, P[i], R[i]);
}

static int cg_check(int64_t N, int L, const void *a, const void *b, const void *scratch)
{
    if (N < 1 || L < 1 || L > CG_MAX_COLUMNS) return fail(SGP_EINVAL, "sgp_cg: need N >= 1 and 1 <= L <= %d", CG_MAX_COLUMNS);
    if (!a || !b || !scratch) return fail(SGP_EINVAL, "sgp_cg: null pointer");
    return SGP_OK;
}

extern "C" size_t sgp_cg_scratch_floats(int L) { return (size_t)CG_MAX_BLOCKS * (size_t)(L > 0 ? L : 1); }

extern "C" int sgp_cg_apply(float *AP, const float *P, const float *s, const float *noise, int64_t N, int L,
                            float *pAp, float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_apply");
    int rc = cg_check(N, L, AP, P, scratch);
    if (rc) return rc;
    if (!s || !noise || !pAp) return fail(SGP_EINVAL, "sgp_cg_apply: null pointer");
    const CgGeometry g = cg_geometry(N, L);
    cudaStream_t st = (cudaStream_t)stream;
    sgp_cg_apply_kernel<<<g.blocks, CG_THREADS, 0, st>>>(AP, P, s, noise, N * (int64_t)L, L, g.active, g.per_block, scratch);
    rc = launch_ok("sgp_cg_apply_kernel");
    if (rc) return rc;
    sgp_cg_reduce_kernel<<<1, CG_THREADS, 0, st>>>(scratch, g.blocks, L, g.active, pAp);
    return launch_ok("sgp_cg_reduce_kernel");
}

extern "C" int sgp_cg_update(float *X, float *R, const float *P, const float *AP, float *rs, const float *pAp,
                             const float *bnorm, float tol, int64_t N, int L, float *alpha_out, float *beta_out,
                             int32_t *done, float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_update");
    int rc = cg_check(N, L, X, R, scratch);
    if (rc) return rc;
    if (!P || !AP || !rs || !pAp || !bnorm || !alpha_out || !beta_out || !done)
        return fail(SGP_EINVAL, "sgp_cg_update: null pointer");
    const CgGeometry g = cg_geometry(N, L);
    cudaStream_t st = (cudaStream_t)stream;
    sgp_cg_update_kernel<<<g.blocks, CG_THREADS, 0, st>>>(X, R, P, AP, rs, pAp, N * (int64_t)L, L, g.active, g.per_block,
                                                          alpha_out, scratch);
    rc = launch_ok("sgp_cg_update_kernel");
    if (rc) return rc;
    sgp_cg_beta_kernel<<<1, CG_THREADS, 0, st>>>(scratch, g.blocks, L, g.active, rs, bnorm, tol, beta_out, done);
    return launch_ok("sgp_cg_beta_kernel");
}

extern "C" int sgp_cg_direction(float *P, const float *R, const float *beta, int64_t N, int L, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_direction");
    if (N < 1 || L < 1 || L > CG_MAX_COLUMNS || !P || !R || !beta) return fail(SGP_EINVAL, "sgp_cg_direction: bad argument");
    const CgGeometry g = cg_geometry(N, L);
    sgp_cg_direction_kernel<<<g.blocks, CG_THREADS, 0, (cudaStream_t)stream>>>(P, R, beta, N * (int64_t)L, L, g.active,
                                                                              g.per_block);
    return launch_ok("sgp_cg_direction_kernel");
}
