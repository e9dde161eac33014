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
    int vec;         // floats per thread and access: 4 (16-byte vectors) when L % 4 == 0 and the blocks are 16-byte aligned, else 1
    int active;      // threads that take part: the largest multiple of L / vec not above CG_THREADS
    int active2;     // threads of the one-block second stage: the largest multiple of L not above CG_THREADS
    int blocks;
    int64_t per_block;   // vec-float elements per block, a multiple of `active`
};

static CgGeometry cg_geometry(int64_t N, int L, bool aligned16)
{
    CgGeometry g;
    g.vec = (L % 4 == 0 && aligned16) ? 4 : 1;
    const int lanes = L / g.vec;
    g.active = (CG_THREADS / lanes) * lanes;
    g.active2 = (CG_THREADS / L) * L;
    const int64_t total = N * (int64_t)lanes;
    int64_t blocks = (total + (int64_t)g.active * 2 * CG_UNROLL - 1) / ((int64_t)g.active * 2 * CG_UNROLL);
    if (blocks > CG_MAX_BLOCKS) blocks = CG_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    int64_t per = (total + blocks - 1) / blocks;
    per = (per + g.active - 1) / g.active * g.active;
    g.blocks = (int)((total + per - 1) / per);
    g.per_block = per;
    return g;
}

// V floats of one row: a thread's unit of work.  V = 4 moves 16-byte vectors (one request per 512 bytes and warp instead
// of four); streaming loads (every block is read once per sweep and is larger than what stays in L2 next to the lattice)
template <int V> struct CgVec {
    float v[V];
};
template <int V> __device__ __forceinline__ CgVec<V> cg_load(const float *p)
{
    CgVec<V> r;
    if (V == 4) {
        const float4 t = __ldcs((const float4 *)p);
        r.v[0] = t.x; r.v[1 % V] = t.y; r.v[2 % V] = t.z; r.v[3 % V] = t.w;
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = __ldcs(p + k);
    }
    return r;
}
template <int V> __device__ __forceinline__ void cg_store(float *p, const CgVec<V> &a)
{
    if (V == 4) {
        *(float4 *)p = make_float4(a.v[0], a.v[1 % V], a.v[2 % V], a.v[3 % V]);
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) p[k] = a.v[k];
    }
}

// sum the per-thread partials of equal column (thread t holds columns (t % (L / V)) * V ... + V - 1) and store them at
// partial[block, :]
template <int V>
__device__ __forceinline__ void cg_block_columns(const float (&acc)[V], int active, int L, float *__restrict__ partial)
{
    __shared__ float s_acc[V][CG_THREADS];
#pragma unroll
    for (int k = 0; k < V; ++k) s_acc[k][threadIdx.x] = threadIdx.x < active ? acc[k] : 0.0f;
    __syncthreads();
    if (threadIdx.x < L) {
        const int lanes = L / V, g = threadIdx.x / V, c = threadIdx.x % V;
        float t = 0.0f;
        for (int k = g; k < active; k += lanes) t += s_acc[c][k];
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

template <int V>
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_apply_kernel(float *__restrict__ AP, const float *__restrict__ P, const float *__restrict__ s_ptr,
                    const float *__restrict__ noise_ptr, int64_t total, int L, int active, int64_t per_block,
                    float *__restrict__ partial)
{
    // total, per_block and the indices below count V-float elements
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.0f;
    if (threadIdx.x < active) {
        const float s = __ldg(s_ptr), noise = __ldg(noise_ptr);
        const int64_t lo = (int64_t)blockIdx.x * per_block;
        const int64_t hi = min(lo + per_block, total);
        int64_t i = lo + threadIdx.x;
        for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
            CgVec<V> p[CG_UNROLL], kp[CG_UNROLL];
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
                const int64_t q = (i + (int64_t)u * active) * V;
                p[u] = cg_load<V>(P + q); kp[u] = cg_load<V>(AP + q);
            }
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    kp[u].v[k] = fmaf(s, kp[u].v[k], noise * p[u].v[k]);
                    acc[k] = fmaf(p[u].v[k], kp[u].v[k], acc[k]);
                }
                cg_store<V>(AP + (i + (int64_t)u * active) * V, kp[u]);
            }
        }
        for (; i < hi; i += active) {
            const CgVec<V> p = cg_load<V>(P + i * V);
            CgVec<V> kp = cg_load<V>(AP + i * V);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                kp.v[k] = fmaf(s, kp.v[k], noise * p.v[k]);
                acc[k] = fmaf(p.v[k], kp.v[k], acc[k]);
            }
            cg_store<V>(AP + i * V, kp);
        }
    }
    cg_block_columns<V>(acc, active, L, partial);
}

__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_reduce_kernel(const float *__restrict__ partial, int blocks, int L, int active, float *__restrict__ out)
{
    const float t = cg_sum_partials(partial, blocks, L, active);
    if (threadIdx.x < L) out[threadIdx.x] = t;
}

template <int V>
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_update_kernel(float *__restrict__ X, float *__restrict__ R, const float *__restrict__ P,
                     const float *__restrict__ AP, const float *__restrict__ rs, const float *__restrict__ pAp,
                     int64_t total, int L, int active, int64_t per_block, float *__restrict__ alpha_out,
                     float *__restrict__ partial)
{
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.0f;
    if (threadIdx.x < active) {
        const int col = (threadIdx.x % (L / V)) * V;
        float alpha[V];
#pragma unroll
        for (int k = 0; k < V; ++k) alpha[k] = __ldg(rs + col + k) / fmaxf(__ldg(pAp + col + k), 1e-30f);
        if (blockIdx.x == 0 && threadIdx.x < L / V) {
#pragma unroll
            for (int k = 0; k < V; ++k) alpha_out[col + k] = alpha[k];
        }
        const int64_t lo = (int64_t)blockIdx.x * per_block;
        const int64_t hi = min(lo + per_block, total);
        int64_t i = lo + threadIdx.x;
        // V = 4: two elements per trip (8 vector loads in flight per thread, as many bytes as the scalar form's 16 x 2)
        constexpr int U = V == 4 ? 2 : CG_UNROLL;
        for (; i + (int64_t)(U - 1) * active < hi; i += (int64_t)U * active) {
            CgVec<V> xv[U], rv[U], pv[U], av[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t q = (i + (int64_t)u * active) * V;
                xv[u] = cg_load<V>(X + q); rv[u] = cg_load<V>(R + q); pv[u] = cg_load<V>(P + q); av[u] = cg_load<V>(AP + q);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t q = (i + (int64_t)u * active) * V;
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    xv[u].v[k] = fmaf(alpha[k], pv[u].v[k], xv[u].v[k]);
                    rv[u].v[k] = fmaf(-alpha[k], av[u].v[k], rv[u].v[k]);
                    acc[k] = fmaf(rv[u].v[k], rv[u].v[k], acc[k]);
                }
                cg_store<V>(X + q, xv[u]);
                cg_store<V>(R + q, rv[u]);
            }
        }
        for (; i < hi; i += active) {
            CgVec<V> xv = cg_load<V>(X + i * V), rv = cg_load<V>(R + i * V);
            const CgVec<V> pv = cg_load<V>(P + i * V), av = cg_load<V>(AP + i * V);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                xv.v[k] = fmaf(alpha[k], pv.v[k], xv.v[k]);
                rv.v[k] = fmaf(-alpha[k], av.v[k], rv.v[k]);
                acc[k] = fmaf(rv.v[k], rv.v[k], acc[k]);
            }
            cg_store<V>(X + i * V, xv);
            cg_store<V>(R + i * V, rv);
        }
    }
    cg_block_columns<V>(acc, active, L, partial);
}

// one block: rs_new from the partials, beta, the convergence flag; rs <- rs_new.
// criterion 0: every column's relative residual sqrt(rs_new) / bnorm is below tol; 1: their mean over the columns with a
// non-zero right-hand side is (GPyTorch's linear_cg stops on residual_norm.mean() < tolerance)
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_beta_kernel(const float *__restrict__ partial, int blocks, int L, int active, float *__restrict__ rs,
                   const float *__restrict__ bnorm, float tol, int criterion, float *__restrict__ beta_out,
                   int32_t *__restrict__ done)
{
    __shared__ int s_open;
    __shared__ float s_rel[CG_THREADS];
    if (threadIdx.x == 0) s_open = 0;
    const float rs_new = cg_sum_partials(partial, blocks, L, active);   // contains a __syncthreads
    __syncthreads();
    if (threadIdx.x < L) {
        beta_out[threadIdx.x] = rs_new / fmaxf(rs[threadIdx.x], 1e-30f);
        rs[threadIdx.x] = rs_new;
        const float rel = sqrtf(rs_new) / bnorm[threadIdx.x];
        s_rel[threadIdx.x] = rel;
        if (criterion == 0 && !(rel < tol)) atomicAdd(&s_open, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (criterion == 0) {
            *done = s_open == 0 ? 1 : 0;
        } else {
            float sum = 0.0f;
            int cols = 0;
            for (int l = 0; l < L; ++l)
                if (bnorm[l] > 1e-29f) { sum += s_rel[l]; ++cols; }   // (bnorm is clamped to 1e-30 for zero columns)
            *done = (cols == 0 || sum / (float)cols < tol) ? 1 : 0;
        }
    }
}

template <int V>
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_direction_kernel(float *__restrict__ P, const float *__restrict__ R, const float *__restrict__ beta,
                        int64_t total, int L, int active, int64_t per_block)
{
    if (threadIdx.x >= active) return;
    const int col = (threadIdx.x % (L / V)) * V;
    float b[V];
#pragma unroll
    for (int k = 0; k < V; ++k) b[k] = __ldg(beta + col + k);
    const int64_t lo = (int64_t)blockIdx.x * per_block;
    const int64_t hi = min(lo + per_block, total);
    int64_t i = lo + threadIdx.x;
    for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
        CgVec<V> pv[CG_UNROLL], rv[CG_UNROLL];
#pragma unroll
        for (int u = 0; u < CG_UNROLL; ++u) {
            const int64_t q = (i + (int64_t)u * active) * V;
            pv[u] = cg_load<V>(P + q); rv[u] = cg_load<V>(R + q);
        }
#pragma unroll
        for (int u = 0; u < CG_UNROLL; ++u) {
#pragma unroll
            for (int k = 0; k < V; ++k) pv[u].v[k] = fmaf(b[k], pv[u].v[k], rv[u].v[k]);
            cg_store<V>(P + (i + (int64_t)u * active) * V, pv[u]);
        }
    }
    for (; i < hi; i += active) {
        CgVec<V> pv = cg_load<V>(P + i * V);
        const CgVec<V> rv = cg_load<V>(R + i * V);
#pragma unroll
        for (int k = 0; k < V; ++k) pv.v[k] = fmaf(b[k], pv.v[k], rv.v[k]);
        cg_store<V>(P + i * V, pv);
    }
}

// ---- the same iteration with the update of X deferred into the direction sweep -------------------------------------
// X += alpha P reads P, which the direction sweep P <- R + beta P reads anyway: moving it there makes the update sweep
// three passes over [N, L] (read R, AP; write R) and the direction sweep five (read R, P, X; write P, X) -- eight
// instead of nine per iteration, same arithmetic in the same order.  The caller applies the last X += alpha P itself
// when the iteration stops after an update.
template <int V>
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_update_r_kernel(float *__restrict__ R, const float *__restrict__ AP, const float *__restrict__ rs,
                       const float *__restrict__ pAp, int64_t total, int L, int active, int64_t per_block,
                       float *__restrict__ alpha_out, float *__restrict__ partial)
{
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.0f;
    if (threadIdx.x < active) {
        const int col = (threadIdx.x % (L / V)) * V;
        float alpha[V];
#pragma unroll
        for (int k = 0; k < V; ++k) alpha[k] = __ldg(rs + col + k) / fmaxf(__ldg(pAp + col + k), 1e-30f);
        if (blockIdx.x == 0 && threadIdx.x < L / V) {
#pragma unroll
            for (int k = 0; k < V; ++k) alpha_out[col + k] = alpha[k];
        }
        const int64_t lo = (int64_t)blockIdx.x * per_block;
        const int64_t hi = min(lo + per_block, total);
        int64_t i = lo + threadIdx.x;
        for (; i + (int64_t)(CG_UNROLL - 1) * active < hi; i += (int64_t)CG_UNROLL * active) {
            CgVec<V> rv[CG_UNROLL], av[CG_UNROLL];
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
                const int64_t q = (i + (int64_t)u * active) * V;
                rv[u] = cg_load<V>(R + q); av[u] = cg_load<V>(AP + q);
            }
#pragma unroll
            for (int u = 0; u < CG_UNROLL; ++u) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    rv[u].v[k] = fmaf(-alpha[k], av[u].v[k], rv[u].v[k]);
                    acc[k] = fmaf(rv[u].v[k], rv[u].v[k], acc[k]);
                }
                cg_store<V>(R + (i + (int64_t)u * active) * V, rv[u]);
            }
        }
        for (; i < hi; i += active) {
            CgVec<V> rv = cg_load<V>(R + i * V);
            const CgVec<V> av = cg_load<V>(AP + i * V);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                rv.v[k] = fmaf(-alpha[k], av.v[k], rv.v[k]);
                acc[k] = fmaf(rv.v[k], rv.v[k], acc[k]);
            }
            cg_store<V>(R + i * V, rv);
        }
    }
    cg_block_columns<V>(acc, active, L, partial);
}

template <int V>
__global__ void __launch_bounds__(CG_THREADS)
sgp_cg_direction_x_kernel(float *__restrict__ P, const float *__restrict__ R, float *__restrict__ X,
                          const float *__restrict__ alpha, const float *__restrict__ beta, int64_t total, int L,
                          int active, int64_t per_block)
{
    if (threadIdx.x >= active) return;
    const int col = (threadIdx.x % (L / V)) * V;
    float a[V], b[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { a[k] = __ldg(alpha + col + k); b[k] = __ldg(beta + col + k); }
    const int64_t lo = (int64_t)blockIdx.x * per_block;
    const int64_t hi = min(lo + per_block, total);
    int64_t i = lo + threadIdx.x;
    constexpr int U = V == 4 ? 2 : CG_UNROLL;
    for (; i + (int64_t)(U - 1) * active < hi; i += (int64_t)U * active) {
        CgVec<V> pv[U], rv[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = (i + (int64_t)u * active) * V;
            pv[u] = cg_load<V>(P + q); rv[u] = cg_load<V>(R + q); xv[u] = cg_load<V>(X + q);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = (i + (int64_t)u * active) * V;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                xv[u].v[k] = fmaf(a[k], pv[u].v[k], xv[u].v[k]);
                pv[u].v[k] = fmaf(b[k], pv[u].v[k], rv[u].v[k]);
            }
            cg_store<V>(X + q, xv[u]);
            cg_store<V>(P + q, pv[u]);
        }
    }
    for (; i < hi; i += active) {
        CgVec<V> pv = cg_load<V>(P + i * V), xv = cg_load<V>(X + i * V);
        const CgVec<V> rv = cg_load<V>(R + i * V);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            xv.v[k] = fmaf(a[k], pv.v[k], xv.v[k]);
            pv.v[k] = fmaf(b[k], pv.v[k], rv.v[k]);
        }
        cg_store<V>(X + i * V, xv);
        cg_store<V>(P + i * V, pv);
    }
}

static bool cg_aligned16(const void *a, const void *b) { return (((uintptr_t)a | (uintptr_t)b) & 15) == 0; }

static int cg_check(int64_t N, int L, const void *a, const void *b, const void *scratch)
{
    if (N < 1 || L < 1 || L > CG_MAX_COLUMNS) return fail(SGP_EINVAL, "sgp_cg: need N >= 1 and 1 <= L <= %d", CG_MAX_COLUMNS);
    if (!a || !b || !scratch) return fail(SGP_EINVAL, "sgp_cg: null pointer");
    return SGP_OK;
}

extern "C" size_t sgp_cg_scratch_floats(int L) { return (size_t)CG_MAX_BLOCKS * (size_t)(L > 0 ? L : 1); }

// second stage alone: out[l] = sum_b partial[b, l] in a fixed order (used by the slice's CG epilogue, sgp_ring.cu)
int sgp_cg_reduce_partials(const float *partial, int blocks, int L, float *out, sgp_stream_t stream)
{
    if (!partial || !out || blocks < 1 || L < 1 || L > CG_MAX_COLUMNS) return fail(SGP_EINVAL, "sgp_cg_reduce_partials: bad argument");
    sgp_cg_reduce_kernel<<<1, CG_THREADS, 0, (cudaStream_t)stream>>>(partial, blocks, L, (CG_THREADS / L) * L, out);
    return launch_ok("sgp_cg_reduce_kernel");
}

extern "C" int sgp_cg_apply(float *AP, const float *P, const float *s, const float *noise, int64_t N, int L,
                            float *pAp, float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_apply");
    int rc = cg_check(N, L, AP, P, scratch);
    if (rc) return rc;
    if (!s || !noise || !pAp) return fail(SGP_EINVAL, "sgp_cg_apply: null pointer");
    const CgGeometry g = cg_geometry(N, L, cg_aligned16(AP, P));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (int64_t)(L / g.vec);
    if (g.vec == 4) sgp_cg_apply_kernel<4><<<g.blocks, CG_THREADS, 0, st>>>(AP, P, s, noise, total, L, g.active, g.per_block, scratch);
    else sgp_cg_apply_kernel<1><<<g.blocks, CG_THREADS, 0, st>>>(AP, P, s, noise, total, L, g.active, g.per_block, scratch);
    rc = launch_ok("sgp_cg_apply_kernel");
    if (rc) return rc;
    sgp_cg_reduce_kernel<<<1, CG_THREADS, 0, st>>>(scratch, g.blocks, L, g.active2, pAp);
    return launch_ok("sgp_cg_reduce_kernel");
}

extern "C" int sgp_cg_update(float *X, float *R, const float *P, const float *AP, float *rs, const float *pAp,
                             const float *bnorm, float tol, int64_t N, int L, float *alpha_out, float *beta_out,
                             int32_t *done, float *scratch, sgp_stream_t stream)
{
    return sgp_cg_update_ex(X, R, P, AP, rs, pAp, bnorm, tol, SGP_CG_ALL_COLUMNS, N, L, alpha_out, beta_out, done, scratch, stream);
}

extern "C" int sgp_cg_update_ex(float *X, float *R, const float *P, const float *AP, float *rs, const float *pAp,
                                const float *bnorm, float tol, int criterion, int64_t N, int L, float *alpha_out,
                                float *beta_out, int32_t *done, float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_update");
    int rc = cg_check(N, L, X, R, scratch);
    if (rc) return rc;
    if (!P || !AP || !rs || !pAp || !bnorm || !alpha_out || !beta_out || !done)
        return fail(SGP_EINVAL, "sgp_cg_update: null pointer");
    if (criterion != SGP_CG_ALL_COLUMNS && criterion != SGP_CG_MEAN) return fail(SGP_EINVAL, "sgp_cg_update: unknown criterion %d", criterion);
    const CgGeometry g = cg_geometry(N, L, cg_aligned16(X, R) && cg_aligned16(P, AP));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (int64_t)(L / g.vec);
    if (g.vec == 4)
        sgp_cg_update_kernel<4><<<g.blocks, CG_THREADS, 0, st>>>(X, R, P, AP, rs, pAp, total, L, g.active, g.per_block, alpha_out, scratch);
    else
        sgp_cg_update_kernel<1><<<g.blocks, CG_THREADS, 0, st>>>(X, R, P, AP, rs, pAp, total, L, g.active, g.per_block, alpha_out, scratch);
    rc = launch_ok("sgp_cg_update_kernel");
    if (rc) return rc;
    sgp_cg_beta_kernel<<<1, CG_THREADS, 0, st>>>(scratch, g.blocks, L, g.active2, rs, bnorm, tol, criterion, beta_out, done);
    return launch_ok("sgp_cg_beta_kernel");
}

extern "C" int sgp_cg_direction(float *P, const float *R, const float *beta, int64_t N, int L, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_direction");
    if (N < 1 || L < 1 || L > CG_MAX_COLUMNS || !P || !R || !beta) return fail(SGP_EINVAL, "sgp_cg_direction: bad argument");
    const CgGeometry g = cg_geometry(N, L, cg_aligned16(P, R));
    const int64_t total = N * (int64_t)(L / g.vec);
    if (g.vec == 4) sgp_cg_direction_kernel<4><<<g.blocks, CG_THREADS, 0, (cudaStream_t)stream>>>(P, R, beta, total, L, g.active, g.per_block);
    else sgp_cg_direction_kernel<1><<<g.blocks, CG_THREADS, 0, (cudaStream_t)stream>>>(P, R, beta, total, L, g.active, g.per_block);
    return launch_ok("sgp_cg_direction_kernel");
}

extern "C" int sgp_cg_update_r(float *R, const float *AP, float *rs, const float *pAp, const float *bnorm, float tol,
                               int criterion, int64_t N, int L, float *alpha_out, float *beta_out, int32_t *done,
                               float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_update_r");
    int rc = cg_check(N, L, R, AP, scratch);
    if (rc) return rc;
    if (!rs || !pAp || !bnorm || !alpha_out || !beta_out || !done) return fail(SGP_EINVAL, "sgp_cg_update_r: null pointer");
    if (criterion != SGP_CG_ALL_COLUMNS && criterion != SGP_CG_MEAN) return fail(SGP_EINVAL, "sgp_cg_update_r: unknown criterion %d", criterion);
    const CgGeometry g = cg_geometry(N, L, cg_aligned16(R, AP));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (int64_t)(L / g.vec);
    if (g.vec == 4) sgp_cg_update_r_kernel<4><<<g.blocks, CG_THREADS, 0, st>>>(R, AP, rs, pAp, total, L, g.active, g.per_block, alpha_out, scratch);
    else sgp_cg_update_r_kernel<1><<<g.blocks, CG_THREADS, 0, st>>>(R, AP, rs, pAp, total, L, g.active, g.per_block, alpha_out, scratch);
    rc = launch_ok("sgp_cg_update_r_kernel");
    if (rc) return rc;
    sgp_cg_beta_kernel<<<1, CG_THREADS, 0, st>>>(scratch, g.blocks, L, g.active2, rs, bnorm, tol, criterion, beta_out, done);
    return launch_ok("sgp_cg_beta_kernel");
}

extern "C" int sgp_cg_direction_x(float *P, const float *R, float *X, const float *alpha, const float *beta, int64_t N,
                                  int L, sgp_stream_t stream)
{
    SGP_RANGE("sgp_cg_direction_x");
    if (N < 1 || L < 1 || L > CG_MAX_COLUMNS || !P || !R || !X || !alpha || !beta)
        return fail(SGP_EINVAL, "sgp_cg_direction_x: bad argument");
    const CgGeometry g = cg_geometry(N, L, cg_aligned16(P, R) && cg_aligned16(X, X));
    const int64_t total = N * (int64_t)(L / g.vec);
    if (g.vec == 4) sgp_cg_direction_x_kernel<4><<<g.blocks, CG_THREADS, 0, (cudaStream_t)stream>>>(P, R, X, alpha, beta, total, L, g.active, g.per_block);
    else sgp_cg_direction_x_kernel<1><<<g.blocks, CG_THREADS, 0, (cudaStream_t)stream>>>(P, R, X, alpha, beta, total, L, g.active, g.per_block);
    return launch_ok("sgp_cg_direction_x_kernel");
}
