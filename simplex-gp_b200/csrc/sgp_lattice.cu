// sgp_lattice.cu -- B200 (sm_100a) kernels and C ABI of the Simplex-GP lattice filter.
//
// Stages (reference = gpytorch_lattice_kernel/cpp/permutohedral.h):
//   build_points      a4-a7  elevate / nearest remainder-0 point / rank / barycentric   (:397-465)
//   hash_insert       a8     lock-free find-or-create of vertex keys                    (:467-474, :58-94)
//   count/number      a8     first-touch numbering of lattice points                    (:73-79)
//   build_neighbours  a9     neighbour-index table of the blur stencil                  (:539-545)
//   splat/blur/slice  a8-a10 the MVM on the built lattice                               (:478-479, :526-556, :497-510)
//
// Arithmetic that decides lattice structure or values is written with the explicit
// round-to-nearest intrinsics (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn), which are
// never contracted into FMAs, in the reference's association order; the reference's
// x86-64 build has no FMA either, so structure is bit-exact and the deterministic
// value paths (gather splat, blur, slice) reproduce the reference's fp32 results.
//
// Everything here is memory/atomic bound (no tensor cores): see DESIGN.md for the
// per-kernel byte counts.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sgp_lattice.h"
#include "sgp_common.cuh"

// ------------------------------------------------------------------------------------
// error plumbing (declared in sgp_common.cuh, shared with sgp_tiles.cu)
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int sgp_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sgp_launch_ok(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sgp_fail(SGP_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SGP_OK;
}

int sgp_pdl_enabled(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SGP_PDL");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v;
}

#define fail sgp_fail
#define launch_ok sgp_launch_ok

#include <nvtx3/nvToolsExt.h>   // header-only: resolves the profiler's injection library at run time, no link dependency
int sgp_nvtx_enabled(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SGP_NVTX");
        v = e ? (atoi(e) != 0) : 0;
    }
    return v;
}
void sgp_nvtx_push(const char *name) { nvtxRangePushA(name); }
void sgp_nvtx_pop(void) { nvtxRangePop(); }

extern "C" int sgp_abi_version(void) { return SGP_ABI_VERSION; }
extern "C" const char *sgp_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------
// host-side constants (same C expressions as the reference, fp32, no contraction)
// ------------------------------------------------------------------------------------
extern "C" int sgp_stencil_variance(const float *coeffs, int k, float *var_out)
{
    if (!coeffs || !var_out || k < 1 || (k & 1) == 0) return fail(SGP_EINVAL, "stencil must have odd length >= 1");
    volatile float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;  // volatile: keep every rounding step
    for (int i = 0; i < k; ++i) {
        float c = coeffs[i];
        m0 = m0 + c;
        volatile float a = (float)i * c;
        m1 = m1 + a;
        volatile float b = (float)(i * i) * c;
        m2 = m2 + b;
    }
    volatile float mean = m1 / m0;
    volatile float msq = mean * mean;
    volatile float q = m2 / m0;
    *var_out = q - msq;
    return SGP_OK;
}

extern "C" int sgp_scale_factors(int d, float var, float *scale_out)
{
    if (d < 1 || d > SGP_MAX_DIM || !scale_out) return fail(SGP_EINVAL, "d must be in [1, %d]", SGP_MAX_DIM);
    for (int i = 0; i < d; ++i) {
        volatile float a = (float)(i + 1) * (float)(i + 2);
        volatile float s = 1.0f / sqrtf(a);
        volatile float v = var + 1.0f / 6.0f;
        volatile float stretch = (float)(d + 1) * sqrtf(v);
        scale_out[i] = s * stretch;
    }
    return SGP_OK;
}

extern "C" float sgp_slice_divisor(int d) { return 1.0f + powf(2.0f, (float)(-d)); }

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
struct ScaleParam {
    float s[SGP_MAX_DIM];
};

// canonical simplex coordinate of vertex `rem` for an axis whose rank is rk (permutohedral.h:364-369)
__device__ __forceinline__ int canon(int rk, int rem, int d) { return (rk <= d - rem) ? rem : rem - (d + 1); }

// ------------------------------------------------------------------------------------
// stage 1a: per-point geometry.  One thread per point, state in registers for D <= 32.
// ------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
sgp_points_kernel(const float *__restrict__ x, int64_t N, int d_rt, int64_t ldx, ScaleParam sp,
                  int16_t *__restrict__ greedy, int8_t *__restrict__ rank, int32_t *__restrict__ replay,
                  int32_t *__restrict__ flags)
{
    constexpr int DM = D > 0 ? D + 1 : SGP_MAX_DIM + 1;
    const int d = D > 0 ? D : d_rt;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;

    float e[DM];
    int g[DM];
    int rk[DM];
    float t[DM];
    const float *p = x + n * ldx;

    // elevation (:397-402)
    {
        float pc = p[d - 1];
        e[d] = __fmul_rn(__fmul_rn((float)(-d), pc), sp.s[d - 1]);
#pragma unroll
        for (int i = (D > 0 ? D : d) - 1; i > 0; --i) {
            float pl = p[i - 1];
            float t1 = __fmul_rn(__fmul_rn((float)i, pl), sp.s[i - 1]);
            float t2 = __fmul_rn(__fmul_rn((float)(i + 2), pc), sp.s[i]);
            e[i] = __fadd_rn(__fsub_rn(e[i + 1], t1), t2);
            pc = pl;
        }
        e[0] = __fadd_rn(e[1], __fmul_rn(__fmul_rn(2.0f, pc), sp.s[0]));
    }

    // nearest remainder-0 point (:404-423)
    const float dp1 = (float)(d + 1);
    const float inv = __fdiv_rn(1.0f, dp1);
    int sum = 0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i <= (D > 0 ? D : d); ++i) {
        float v = __fmul_rn(e[i], inv);
        float up = __fmul_rn(ceilf(v), dp1);
        float dn = __fmul_rn(floorf(v), dp1);
        float pick = (__fsub_rn(up, e[i]) < __fsub_rn(e[i], dn)) ? up : dn;
        bad |= !(pick >= -32768.0f && pick <= 32767.0f);
        g[i] = (int)(int16_t)__float2int_rz(pick);
        sum += g[i];
    }
    sum = __float2int_rz(__fmul_rn((float)sum, inv));

    // rank of the differentials (:427-433), ties to the later index
#pragma unroll
    for (int i = 0; i <= (D > 0 ? D : d); ++i) {
        t[i] = __fsub_rn(e[i], (float)g[i]);
        rk[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < (D > 0 ? D : d); ++i) {
#pragma unroll
        for (int j = i + 1; j <= (D > 0 ? D : d); ++j) {
            if (t[i] < t[j]) rk[i]++; else rk[j]++;
        }
    }

    // back onto the hyperplane (:435-457)
    if (sum > 0) {
#pragma unroll
        for (int i = 0; i <= (D > 0 ? D : d); ++i) {
            if (rk[i] >= d + 1 - sum) {
                g[i] = (int)(int16_t)(g[i] - (d + 1));
                rk[i] += sum - (d + 1);
            } else {
                rk[i] += sum;
            }
        }
    } else if (sum < 0) {
#pragma unroll
        for (int i = 0; i <= (D > 0 ? D : d); ++i) {
            if (rk[i] < -sum) {
                g[i] = (int)(int16_t)(g[i] + (d + 1));
                rk[i] += (d + 1) + sum;
            } else {
                rk[i] += sum;
            }
        }
    }

    // barycentric weights (:459-465).  With s[q] = the scaled differential of the axis whose
    // rank is q, the reference's accumulation gives b[k] = s[d-k] - s[d+1-k] (1<=k<=d) and
    // b[0] = s[d] + (1 - s[0]), whatever the visiting order (one add and one subtract per slot).
#pragma unroll
    for (int i = 0; i <= (D > 0 ? D : d); ++i) t[i] = __fmul_rn(__fsub_rn(e[i], (float)g[i]), inv);
    float s[DM];
    if (D > 0) {
#pragma unroll
        for (int q = 0; q <= (D > 0 ? D : 0); ++q) {
            float v = 0.0f;
#pragma unroll
            for (int i = 0; i <= (D > 0 ? D : 0); ++i) v = (rk[i] == q) ? t[i] : v;
            s[q] = v;
        }
    } else {
        for (int i = 0; i <= d; ++i) s[i] = 0.0f;
        for (int i = 0; i <= d; ++i) {
            int q = rk[i];
            if (q >= 0 && q <= d) s[q] = t[i];
        }
    }

    const int64_t base = n * (d + 1);
#pragma unroll
    for (int k = 0; k <= (D > 0 ? D : d); ++k) {
        float w = (k == 0) ? __fadd_rn(s[d], __fsub_rn(1.0f, s[0])) : __fsub_rn(s[d - k], s[d + 1 - k]);
        replay[(base + k) * 2 + 1] = __float_as_int(w);
        greedy[base + k] = (int16_t)g[k];
        rank[base + k] = (int8_t)rk[k];
    }
    if (bad) atomicOr(flags, SGP_FLAG_KEY_RANGE);
}

#define SGP_MAX_PROBES 8192ull

// ------------------------------------------------------------------------------------
// stage 1b: lock-free hash insertion.  One thread per point-vertex pv = n*(d+1)+rem.
// A slot is a 64-bit word {fingerprint:32 | owner pv:32}; it is claimed with one CAS and
// afterwards only ever lowered (atomicMin) to a smaller pv carrying the same key, so the
// final owner of every key is its first toucher in the reference's sequential order.
// Keys are never stored in the table: the owner's key is recomputed from greedy/rank,
// which a previous launch wrote.
// ------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
sgp_insert_kernel(const int16_t *__restrict__ greedy, const int8_t *__restrict__ rank, int64_t N, int d_rt,
                  unsigned long long *table, uint64_t mask, uint32_t *__restrict__ slot_of,
                  int32_t *__restrict__ flags)
{
    constexpr int DK = D > 0 ? D : SGP_MAX_DIM;
    const int d = D > 0 ? D : d_rt;
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = N * (d + 1);
    if (pv >= total) return;
    const int64_t n = pv / (d + 1);
    const int rem = (int)(pv - n * (d + 1));

    int16_t key[DK];
    {
        const int16_t *gp = greedy + n * (d + 1);
        const int8_t *rp = rank + n * (d + 1);
#pragma unroll
        for (int i = 0; i < (D > 0 ? D : d); ++i) key[i] = (int16_t)(gp[i] + canon(rp[i], rem, d));
    }
    const uint64_t h = hash_key<D>(key, d);
    const uint32_t fp = (uint32_t)(h >> 32);
    uint64_t slot = h & mask;
    const unsigned long long mine = ((unsigned long long)fp << 32) | (uint32_t)pv;

    // A caller may size the table from the previous lattice of the same shape (4x its M instead of 2 N(d+1)); if this
    // lattice outgrew it, give up quickly instead of scanning a full table from every thread: a probe sequence longer
    // than SGP_MAX_PROBES never occurs below ~95 % load, and once one thread has raised the flag the others leave at
    // their next check.
    const uint64_t max_probes = mask + 1 < SGP_MAX_PROBES ? mask + 1 : SGP_MAX_PROBES;
    for (uint64_t probe = 0; probe < max_probes; ++probe) {
        if ((probe & 63) == 63 && (*((volatile int32_t *)flags) & SGP_FLAG_TABLE_FULL)) break;
        unsigned long long cur = *((volatile unsigned long long *)(table + slot));
        if (cur == SGP_EMPTY) {
            unsigned long long prev = atomicCAS(table + slot, SGP_EMPTY, mine);
            if (prev == SGP_EMPTY) {
                slot_of[pv] = (uint32_t)slot;
                return;
            }
            cur = prev;
        }
        if ((uint32_t)(cur >> 32) == fp) {
            const uint32_t opv = (uint32_t)cur;
            const int64_t on = opv / (uint32_t)(d + 1);
            const int orem = (int)(opv - on * (d + 1));
            const int16_t *gp = greedy + on * (d + 1);
            const int8_t *rp = rank + on * (d + 1);
            bool same = true;
#pragma unroll
            for (int i = 0; i < (D > 0 ? D : d); ++i) {
                int16_t ok = (int16_t)(gp[i] + canon(rp[i], orem, d));
                same = same && (ok == key[i]);
            }
            if (same) {
                if ((uint32_t)pv < opv) atomicMin(table + slot, mine);
                slot_of[pv] = (uint32_t)slot;
                return;
            }
        }
        slot = (slot + 1) & mask;
    }
    atomicOr(flags, SGP_FLAG_TABLE_FULL);
    slot_of[pv] = 0;
}

// ------------------------------------------------------------------------------------
// stage 1c: first-touch numbering.  mark -> exclusive scan -> renumber.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgp_mark_kernel(const unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of,
                int64_t total, uint32_t *__restrict__ marks)
{
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= total) return;
    const unsigned long long e = table[slot_of[pv]];
    marks[pv] = ((uint32_t)e == (uint32_t)pv) ? 1u : 0u;
}

// Three-launch exclusive scan of uint32 (tile = 256 threads x 16 items).
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total_out)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t warp_off = wid > 0 ? warp_sums[wid - 1] : 0u;
    if (total_out) *total_out = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
sgp_scan_reduce_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += in[i];
    }
    uint32_t total;
    block_exclusive_scan(acc, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place; grand total -> *total_out (uint64)
__global__ void __launch_bounds__(SCAN_THREADS)
sgp_scan_spine_kernel(uint32_t *__restrict__ tile_sums, int64_t n_tiles, unsigned long long *__restrict__ total_out)
{
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        int64_t i = base + threadIdx.x;
        uint32_t v = (i < n_tiles) ? tile_sums[i] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        unsigned long long carry = carry_s;
        if (i < n_tiles) tile_sums[i] = (uint32_t)(carry + ex);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS)
sgp_scan_down_kernel(uint32_t *__restrict__ data, int64_t n, const uint32_t *__restrict__ tile_offsets)
{
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        v[k] = (i < n) ? data[i] : 0u;
        acc += v[k];
    }
    uint32_t ex = block_exclusive_scan(acc, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n) data[i] = ex;
        ex += v[k];
    }
}

// exclusive scan of data[n] in place; tile_sums: scratch of ceil(n/SCAN_TILE) uint32; total -> device uint64
int sgp_exclusive_scan_u32(uint32_t *data, int64_t n, uint32_t *tile_sums, unsigned long long *total_dev,
                             cudaStream_t st)
{
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles == 0) {
        CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), st));
        return SGP_OK;
    }
    sgp_scan_reduce_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(data, n, tile_sums);
    sgp_scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(tile_sums, n_tiles, total_dev);
    sgp_scan_down_kernel<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(data, n, tile_sums);
    return launch_ok("exclusive scan");
}

template <int D>
__global__ void __launch_bounds__(256)
sgp_renumber_kernel(const unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of,
                    const uint32_t *__restrict__ pos, const int16_t *__restrict__ greedy,
                    const int8_t *__restrict__ rank, int64_t N, int d_rt, int32_t *__restrict__ replay,
                    int16_t *__restrict__ keys)
{
    const int d = D > 0 ? D : d_rt;
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= N * (d + 1)) return;
    const uint32_t owner = (uint32_t)table[slot_of[pv]];
    const uint32_t idx = pos[owner];
    replay[pv * 2] = (int32_t)idx;
    if (owner == (uint32_t)pv) {
        const int64_t n = pv / (d + 1);
        const int rem = (int)(pv - n * (d + 1));
        const int16_t *gp = greedy + n * (d + 1);
        const int8_t *rp = rank + n * (d + 1);
        int16_t *kp = keys + (int64_t)idx * d;
#pragma unroll
        for (int i = 0; i < (D > 0 ? D : d); ++i) kp[i] = (int16_t)(gp[i] + canon(rp[i], rem, d));
    }
}

// table slot {fp | owner pv} -> {fp | lattice index}
__global__ void __launch_bounds__(256)
sgp_retarget_kernel(unsigned long long *__restrict__ table, int64_t capacity, const uint32_t *__restrict__ pos)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= capacity) return;
    const unsigned long long e = table[s];
    if (e == SGP_EMPTY) return;
    table[s] = (e & 0xFFFFFFFF00000000ull) | pos[(uint32_t)e];
}

// ------------------------------------------------------------------------------------
// Extension of a built lattice by more points (the union lattice of the rectangular operator,
// bilateral_kernel.py:142-160).  First-touch numbering is sequential over the points, so the lattice of
// cat([x_old, x_new]) numbers the old keys exactly as the lattice of x_old alone and appends the keys only the new
// points touch, in their own first-touch order.  The table is therefore seeded from keys[M_old] ({fp | index}), the new
// point-vertices are inserted against it, and only they are marked, scanned and numbered.  A slot claimed by a new
// point-vertex carries SGP_NEW_OWNER in its low word until it is numbered.
// ------------------------------------------------------------------------------------
#define SGP_NEW_OWNER 0x80000000u

template <int D>
__global__ void __launch_bounds__(256)
sgp_seed_kernel(const int16_t *__restrict__ keys, int64_t M, int d_rt, unsigned long long *table, uint64_t mask,
                int32_t *__restrict__ flags)
{
    constexpr int DK = D > 0 ? D : SGP_MAX_DIM;
    const int d = D > 0 ? D : d_rt;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int16_t key[DK];
#pragma unroll
    for (int c = 0; c < (D > 0 ? D : d); ++c) key[c] = keys[i * d + c];
    const uint64_t h = hash_key<D>(key, d);
    const unsigned long long mine = ((unsigned long long)(uint32_t)(h >> 32) << 32) | (uint32_t)i;
    uint64_t slot = h & mask;
    for (uint64_t probe = 0; probe <= mask; ++probe) {   // the keys are distinct: any empty slot on the probe path will do
        if (atomicCAS(table + slot, SGP_EMPTY, mine) == SGP_EMPTY) return;
        slot = (slot + 1) & mask;
    }
    atomicOr(flags, SGP_FLAG_TABLE_FULL);
}

template <int D>
__global__ void __launch_bounds__(256)
sgp_extend_insert_kernel(const int16_t *__restrict__ greedy, const int8_t *__restrict__ rank, int64_t N, int d_rt,
                         const int16_t *__restrict__ keys, unsigned long long *table, uint64_t mask,
                         uint32_t *__restrict__ slot_of, int32_t *__restrict__ flags)
{
    constexpr int DK = D > 0 ? D : SGP_MAX_DIM;
    const int d = D > 0 ? D : d_rt;
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= N * (d + 1)) return;
    const int64_t n = pv / (d + 1);
    const int rem = (int)(pv - n * (d + 1));
    int16_t key[DK];
    {
        const int16_t *gp = greedy + n * (d + 1);
        const int8_t *rp = rank + n * (d + 1);
#pragma unroll
        for (int i = 0; i < (D > 0 ? D : d); ++i) key[i] = (int16_t)(gp[i] + canon(rp[i], rem, d));
    }
    const uint64_t h = hash_key<D>(key, d);
    const uint32_t fp = (uint32_t)(h >> 32);
    uint64_t slot = h & mask;
    const unsigned long long mine = ((unsigned long long)fp << 32) | SGP_NEW_OWNER | (uint32_t)pv;

    for (uint64_t probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = *((volatile unsigned long long *)(table + slot));
        if (cur == SGP_EMPTY) {
            unsigned long long prev = atomicCAS(table + slot, SGP_EMPTY, mine);
            if (prev == SGP_EMPTY) {
                slot_of[pv] = (uint32_t)slot;
                return;
            }
            cur = prev;
        }
        if ((uint32_t)(cur >> 32) == fp) {
            const uint32_t low = (uint32_t)cur;
            bool same = true;
            if (low & SGP_NEW_OWNER) {     // claimed by another new point-vertex: its key comes from greedy/rank
                const uint32_t opv = low & ~SGP_NEW_OWNER;
                const int64_t on = opv / (uint32_t)(d + 1);
                const int orem = (int)(opv - on * (d + 1));
                const int16_t *gp = greedy + on * (d + 1);
                const int8_t *rp = rank + on * (d + 1);
#pragma unroll
                for (int i = 0; i < (D > 0 ? D : d); ++i)
                    same = same && ((int16_t)(gp[i] + canon(rp[i], orem, d)) == key[i]);
                if (same && (uint32_t)pv < opv) atomicMin(table + slot, mine);
            } else {                       // a lattice point of the lattice being extended
                const int16_t *kp = keys + (int64_t)low * d;
#pragma unroll
                for (int i = 0; i < (D > 0 ? D : d); ++i) same = same && (kp[i] == key[i]);
            }
            if (same) {
                slot_of[pv] = (uint32_t)slot;
                return;
            }
        }
        slot = (slot + 1) & mask;
    }
    atomicOr(flags, SGP_FLAG_TABLE_FULL);
    slot_of[pv] = 0;
}

__global__ void __launch_bounds__(256)
sgp_extend_mark_kernel(const unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of,
                       int64_t total, uint32_t *__restrict__ marks)
{
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= total) return;
    marks[pv] = ((uint32_t)table[slot_of[pv]] == (SGP_NEW_OWNER | (uint32_t)pv)) ? 1u : 0u;
}

template <int D>
__global__ void __launch_bounds__(256)
sgp_extend_renumber_kernel(const unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of,
                           const uint32_t *__restrict__ pos, const int16_t *__restrict__ greedy,
                           const int8_t *__restrict__ rank, int64_t N, int d_rt, uint32_t M_old,
                           int32_t *__restrict__ replay, int16_t *__restrict__ keys)
{
    const int d = D > 0 ? D : d_rt;
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= N * (d + 1)) return;
    const uint32_t low = (uint32_t)table[slot_of[pv]];
    const uint32_t idx = (low & SGP_NEW_OWNER) ? M_old + pos[low & ~SGP_NEW_OWNER] : low;
    replay[pv * 2] = (int32_t)idx;
    if (low == (SGP_NEW_OWNER | (uint32_t)pv)) {
        const int64_t n = pv / (d + 1);
        const int rem = (int)(pv - n * (d + 1));
        const int16_t *gp = greedy + n * (d + 1);
        const int8_t *rp = rank + n * (d + 1);
        int16_t *kp = keys + (int64_t)idx * d;
#pragma unroll
        for (int i = 0; i < (D > 0 ? D : d); ++i) kp[i] = (int16_t)(gp[i] + canon(rp[i], rem, d));
    }
}

// slots owned by new point-vertices: {fp | NEW | pv} -> {fp | lattice index}.  Only the owner writes its slot; a
// reader that sees the rewritten word no longer matches NEW | pv and leaves it alone.
__global__ void __launch_bounds__(256)
sgp_extend_retarget_kernel(unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of, int64_t total,
                           const uint32_t *__restrict__ pos, uint32_t M_old)
{
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= total) return;
    const uint32_t slot = slot_of[pv];
    const unsigned long long e = table[slot];
    if ((uint32_t)e == (SGP_NEW_OWNER | (uint32_t)pv)) table[slot] = (e & 0xFFFFFFFF00000000ull) | (M_old + pos[pv]);
}

// ------------------------------------------------------------------------------------
// Appending an explicit list of DISTINCT keys to a seeded table: the merge step of a point-sharded build.  Rank g
// builds the lattice of its own points; the global first-touch numbering is then the concatenation, in rank order, of
// every rank's keys that no earlier rank holds, each list in its own (local first-touch) order -- ranks own contiguous
// point ranges, so this IS the sequential order of permutohedral.h:73-79,467-485.  Every rank replays that merge on the
// all-gathered key lists and ends with the same keys[M, d] and, for its own list, the local -> global index map.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgp_append_insert_kernel(const int16_t *__restrict__ new_keys, int64_t m_new, int d, const int16_t *__restrict__ keys,
                         unsigned long long *table, uint64_t mask, uint32_t *__restrict__ slot_of,
                         int32_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_new) return;
    const int16_t *key = new_keys + i * d;
    const uint64_t h = hash_key<0>(key, d);
    const uint32_t fp = (uint32_t)(h >> 32);
    uint64_t slot = h & mask;
    const unsigned long long mine = ((unsigned long long)fp << 32) | SGP_NEW_OWNER | (uint32_t)i;
    for (uint64_t probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = *((volatile unsigned long long *)(table + slot));
        if (cur == SGP_EMPTY) {
            const unsigned long long prev = atomicCAS(table + slot, SGP_EMPTY, mine);
            if (prev == SGP_EMPTY) {
                slot_of[i] = (uint32_t)slot;
                return;
            }
            cur = prev;
        }
        if ((uint32_t)(cur >> 32) == fp) {
            const uint32_t low = (uint32_t)cur;
            const int16_t *op = (low & SGP_NEW_OWNER) ? new_keys + (int64_t)(low & ~SGP_NEW_OWNER) * d : keys + (int64_t)low * d;
            bool same = true;
            for (int c = 0; c < d; ++c) same = same && (op[c] == key[c]);
            if (same) {
                if ((low & SGP_NEW_OWNER) && (uint32_t)i < (low & ~SGP_NEW_OWNER)) atomicMin(table + slot, mine);   // duplicate in the list
                slot_of[i] = (uint32_t)slot;
                return;
            }
        }
        slot = (slot + 1) & mask;
    }
    atomicOr(flags, SGP_FLAG_TABLE_FULL);
    slot_of[i] = 0;
}

__global__ void __launch_bounds__(256)
sgp_append_number_kernel(const unsigned long long *__restrict__ table, const uint32_t *__restrict__ slot_of,
                         const uint32_t *__restrict__ pos, const int16_t *__restrict__ new_keys, int64_t m_new, int d,
                         uint32_t M_old, int32_t *__restrict__ map_out, int16_t *__restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_new) return;
    const uint32_t low = (uint32_t)table[slot_of[i]];
    const uint32_t idx = (low & SGP_NEW_OWNER) ? M_old + pos[low & ~SGP_NEW_OWNER] : low;
    if (map_out) map_out[i] = (int32_t)idx;
    if (low == (SGP_NEW_OWNER | (uint32_t)i)) {
        const int16_t *key = new_keys + i * d;
        int16_t *kp = keys + (int64_t)idx * d;
        for (int c = 0; c < d; ++c) kp[c] = key[c];
    }
}

// ------------------------------------------------------------------------------------
// stage 1d: neighbour table.  One thread per (axis j, lattice point i).
// ------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
sgp_neighbours_kernel(const int16_t *__restrict__ keys, int64_t M, int d_rt, int order,
                      const unsigned long long *__restrict__ table, uint64_t mask, int32_t *__restrict__ nbr)
{
    constexpr int DK = D > 0 ? D : SGP_MAX_DIM;
    const int d = D > 0 ? D : d_rt;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= M * (d + 1)) return;
    const int j = (int)(tid / M);
    const int64_t i = tid - (int64_t)j * M;

    int16_t key[DK];
    const int16_t *kp = keys + i * d;
#pragma unroll
    for (int c = 0; c < (D > 0 ? D : d); ++c) key[c] = kp[c];

    int32_t *out = nbr + ((int64_t)j * M + i) * (2 * order);
    int t = 0;
    for (int o = -order; o <= order; ++o) {
        if (o == 0) continue;
        int16_t nk[DK];
#pragma unroll
        for (int c = 0; c < (D > 0 ? D : d); ++c) {
            int v = (int)key[c] - o;
            if (c == j) v = (int)key[c] + o * d;
            nk[c] = (int16_t)v;
        }
        const uint64_t h = hash_key<D>(nk, d);
        const uint32_t fp = (uint32_t)(h >> 32);
        uint64_t slot = h & mask;
        int32_t found = -1;
        for (uint64_t probe = 0; probe <= mask; ++probe) {
            const unsigned long long cur = table[slot];
            if (cur == SGP_EMPTY) break;
            if ((uint32_t)(cur >> 32) == fp) {
                const uint32_t idx = (uint32_t)cur;
                const int16_t *op = keys + (int64_t)idx * d;
                bool same = true;
#pragma unroll
                for (int c = 0; c < (D > 0 ? D : d); ++c) same = same && (op[c] == nk[c]);
                if (same) {
                    found = (int32_t)idx;
                    break;
                }
            }
            slot = (slot + 1) & mask;
        }
        out[t++] = found;
    }
}

// ------------------------------------------------------------------------------------
// stages 2-4: splat / blur / slice.  Thread = (row, chunk of VEC channels); a row's
// channels are adjacent so one row of L=16 fp32 is 4 lanes x float4 = 64 contiguous bytes.
// ------------------------------------------------------------------------------------
#define SLICE_BATCH 9
// splat, scatter form: thread = (point n, chunk); (d+1) vector reductions into the lattice
template <int VEC>
__global__ void __launch_bounds__(256)
sgp_splat_atomic_kernel(const int2 *__restrict__ replay, int64_t pstride, int64_t rstride,
                        const uint32_t *__restrict__ perm, const float *__restrict__ src, int64_t lds, int64_t N,
                        int dp1, int L, int chunks, float *__restrict__ values)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / chunks;
    if (n >= N) return;
    const int c0 = (int)(tid - n * chunks) * VEC;
    Vec<VEC> v;
    v.load(src + (perm ? (int64_t)__ldg(perm + n) : n) * lds + c0);   // replay is in processing order, src in caller order
    // entry (n, r) lives at replay[n*pstride + r*rstride]: [N, d+1] (pstride = d+1, rstride = 1) or the transposed
    // [d+1, N] (pstride = 1, rstride = N), for which the 8 points of a warp read one 64-byte run per vertex
    const int2 *rp = replay + n * pstride;
    for (int r0 = 0; r0 < dp1; r0 += SLICE_BATCH) {
        int2 e[SLICE_BATCH];
#pragma unroll
        for (int b = 0; b < SLICE_BATCH; ++b)
            e[b] = (r0 + b < dp1) ? ldg_ordered_int2(rp + (r0 + b) * rstride) : make_int2(0, 0);
#pragma unroll
        for (int b = 0; b < SLICE_BATCH; ++b) {
            if (r0 + b < dp1) {
                const float w = __int_as_float(e[b].y);
                Vec<VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; ++k) o.v[k] = __fmul_rn(w, v.v[k]);
                o.red(values + (int64_t)e[b].x * L + c0);
            }
        }
    }
}

// splat, gather form: thread = (lattice point i, chunk); sequential sum over the row's
// (point, weight) list in point-vertex order, which is the reference's accumulation order.
template <int VEC>
__global__ void __launch_bounds__(256)
sgp_splat_gather_kernel(const uint32_t *__restrict__ row_ptr, const int2 *__restrict__ entries,
                        const float *__restrict__ src, int64_t lds, int64_t M, int L, int chunks,
                        float *__restrict__ values)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = tid / chunks;
    if (i >= M) return;
    const int c0 = (int)(tid - i * chunks) * VEC;
    const uint32_t a = __ldg(row_ptr + i), b = __ldg(row_ptr + i + 1);
    Vec<VEC> acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
    for (uint32_t q = a; q < b; ++q) {
        const int2 e = __ldg(entries + sgp_entry_index(q));   // row-sorted position -> interleaved storage
        const float w = __int_as_float(e.y);
        Vec<VEC> v;
        v.load(src + (int64_t)(e.x & 0x7fffffff) * lds + c0);   // bit 31 is the row-start flag of the row-sorted entries
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = __fadd_rn(acc.v[k], __fmul_rn(w, v.v[k]));
    }
    acc.store(values + i * L + c0);
}

struct CoeffParam {
    float c[2 * SGP_MAX_ORDER + 1];
};

// one blur pass along axis j: thread = (ROWS lattice points, one chunk).  All neighbour indices of the
// thread's rows are loaded first, then all (2r+1)*ROWS lattice rows, so that several dependent
// index->row load chains overlap (the pass is latency-bound otherwise: M*L/4 threads is only ~5 waves).
template <int VEC, int R, int ROWS, bool FAST>
__global__ void __launch_bounds__(256)
sgp_blur_kernel(const int32_t *__restrict__ nbr_j, const float *__restrict__ in, float *__restrict__ out,
                int64_t M, int L, int chunks, int order_rt, CoeffParam cf)
{
    constexpr int RR = R > 0 ? R : SGP_MAX_ORDER;
    const int r = R > 0 ? R : order_rt;
    const int rows_per_block = blockDim.x / chunks;           // host guarantees blockDim.x % chunks == 0
    const int lr = threadIdx.x / chunks;
    const int c0 = (threadIdx.x - lr * chunks) * VEC;
    const int64_t base = (int64_t)blockIdx.x * rows_per_block * ROWS + lr;

    int32_t nb[ROWS][2 * RR];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const int64_t i = base + (int64_t)q * rows_per_block;
        const int32_t *np = nbr_j + i * (2 * r);
#pragma unroll
        for (int t = 0; t < 2 * RR; ++t) nb[q][t] = (i < M && t < 2 * r) ? __ldg(np + t) : -1;
    }
    Vec<VEC> v[ROWS][2 * RR + 1];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const int64_t i = base + (int64_t)q * rows_per_block;
#pragma unroll
        for (int t = 0; t < 2 * RR; ++t)
            if (t < 2 * r && nb[q][t] >= 0) v[q][t].load(in + (int64_t)nb[q][t] * L + c0);
        if (i < M) v[q][2 * RR].load(in + i * L + c0);
    }
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const int64_t i = base + (int64_t)q * rows_per_block;
        if (i >= M) continue;
        Vec<VEC> acc;
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
        // reference order: o = -r..-1, 0, 1..r
#pragma unroll
        for (int t = 0; t < RR; ++t) {
            if (t < r && nb[q][t] >= 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc.v[k] = madd<FAST>(cf.c[t], v[q][t].v[k], acc.v[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = madd<FAST>(cf.c[r], v[q][2 * RR].v[k], acc.v[k]);
#pragma unroll
        for (int t = 0; t < RR; ++t) {
            if (t < r && nb[q][r + t] >= 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc.v[k] = madd<FAST>(cf.c[r + 1 + t], v[q][r + t].v[k], acc.v[k]);
            }
        }
        acc.store(out + i * L + c0);
    }
}

// counts[0]: mismatches with |a| in [2^-100, inf) or a == 0 (must be 0);
// counts[1]: inputs below 2^-100 whose absolute error exceeds 1e-37 (must be 0).
__global__ void __launch_bounds__(256)
sgp_exact_div_check_kernel(float b, float rb, uint32_t lo, uint32_t count, unsigned long long *counts)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float a = __uint_as_float(lo + (uint32_t)i);
    if (!isfinite(a)) return;
    const float want = __fdiv_rn(a, b);
    const float got = exact_div(a, b, rb);
    const float mag = fabsf(a);
    if (mag >= 7.888609052210118e-31f || mag == 0.0f) {   // 2^-100
        if (want != got) atomicAdd(counts, 1ull);
    } else {
        if (!(fabsf(want - got) <= 1e-37f)) atomicAdd(counts + 1, 1ull);
    }
}

// slice: thread = (point n, chunk).  Vertices are processed in batches of BATCH with all
// replay entries, then all lattice rows, in flight together (two dependent latencies per batch
// instead of two per vertex); the sum itself stays in vertex order.
// RAGGED (only instantiated with STREAM): out has L_out <= L columns and arbitrary row alignment; it is written one
// channel at a time (the lattice rows stay 16-byte vectors).
template <int VEC, int BATCH, bool FAST, bool STREAM, bool RAGGED>
__global__ void __launch_bounds__(256)
sgp_slice_kernel(const int2 *__restrict__ replay, int64_t pstride, int64_t rstride,
                 const uint32_t *__restrict__ perm, const float *__restrict__ values, int64_t N, int dp1, int L,
                 int chunks, float divisor, float rdivisor, float *__restrict__ out, int64_t ldo, int L_out)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / chunks;
    if (n >= N) return;
    const int c0 = (int)(tid - n * chunks) * VEC;
    const int2 *rp = replay + n * pstride;
    pdl_launch_dependents();
    bool waited = false;    // the replay entries do not depend on the blur: they are loaded before the wait
    Vec<VEC> acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] = 0.0f;
    int r0 = 0;
    for (; r0 + BATCH <= dp1; r0 += BATCH) {   // full batches: no predicates
        int2 e[BATCH];
        Vec<VEC> v[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) e[b] = STREAM ? ldg_ordered_int2_streaming(rp + (r0 + b) * rstride) : ldg_ordered_int2(rp + (r0 + b) * rstride);
        // A warp issues in order, so a row load placed before a later index load would make the warp wait a full
        // memory round trip with that index load not yet issued.  Lattice indices are never negative: branching on
        // all of them keeps every index load ahead of every row load (two round trips per batch, not up to BATCH).
        int lowest = e[0].x;
#pragma unroll
        for (int b = 1; b < BATCH; ++b) lowest = min(lowest, e[b].x);
        if (!waited) { pdl_wait(); waited = true; }
        if (lowest >= 0) {
#pragma unroll
            for (int b = 0; b < BATCH; ++b) v[b].load_ordered(values + (int64_t)e[b].x * L + c0);
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const float w = __int_as_float(e[b].y);
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc.v[k] = FAST ? __fmaf_rn(w, v[b].v[k], acc.v[k])
                                    : __fadd_rn(acc.v[k], exact_div(__fmul_rn(w, v[b].v[k]), divisor, rdivisor));
            }
        }
    }
    if (!waited) pdl_wait();
    for (; r0 < dp1; ++r0) {
        const int2 e = __ldg(rp + r0 * rstride);
        const float w = __int_as_float(e.y);
        Vec<VEC> v;
        v.load(values + (int64_t)e.x * L + c0);
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            acc.v[k] = FAST ? __fmaf_rn(w, v.v[k], acc.v[k])
                            : __fadd_rn(acc.v[k], exact_div(__fmul_rn(w, v.v[k]), divisor, rdivisor));
    }
    if (FAST) {   // one division of the sum instead of one per term (differs from the reference by rounding only)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = exact_div(acc.v[k], divisor, rdivisor);
    }
    float *orow = out + (perm ? (int64_t)__ldg(perm + n) : n) * ldo + c0;
    if (RAGGED) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (c0 + k < L_out) __stcs(orow + k, acc.v[k]);
    } else if (STREAM) {
        acc.store_streaming(orow);
    } else {
        acc.store(orow);
    }
}

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t work, int block) { return (unsigned)((work + block - 1) / block); }

#define SGP_DISPATCH_D(d, CALL)                                                             \
    switch (d) {                                                                            \
        case 1: { constexpr int DD = 1; CALL; } break;                                      \
        case 2: { constexpr int DD = 2; CALL; } break;                                      \
        case 3: { constexpr int DD = 3; CALL; } break;                                      \
        case 4: { constexpr int DD = 4; CALL; } break;                                      \
        case 5: { constexpr int DD = 5; CALL; } break;                                      \
        case 6: { constexpr int DD = 6; CALL; } break;                                      \
        case 7: { constexpr int DD = 7; CALL; } break;                                      \
        case 8: { constexpr int DD = 8; CALL; } break;                                      \
        case 9: { constexpr int DD = 9; CALL; } break;                                      \
        case 10: { constexpr int DD = 10; CALL; } break;                                    \
        case 11: { constexpr int DD = 11; CALL; } break;                                    \
        case 12: { constexpr int DD = 12; CALL; } break;                                    \
        case 13: { constexpr int DD = 13; CALL; } break;                                    \
        case 14: { constexpr int DD = 14; CALL; } break;                                    \
        case 15: { constexpr int DD = 15; CALL; } break;                                    \
        case 16: { constexpr int DD = 16; CALL; } break;                                    \
        case 17: { constexpr int DD = 17; CALL; } break;                                    \
        case 18: { constexpr int DD = 18; CALL; } break;                                    \
        case 19: { constexpr int DD = 19; CALL; } break;                                    \
        case 20: { constexpr int DD = 20; CALL; } break;                                    \
        case 21: { constexpr int DD = 21; CALL; } break;                                    \
        case 22: { constexpr int DD = 22; CALL; } break;                                    \
        case 23: { constexpr int DD = 23; CALL; } break;                                    \
        case 24: { constexpr int DD = 24; CALL; } break;                                    \
        default: { constexpr int DD = 0; CALL; } break;                                     \
    }

static int check_dims(int64_t N, int d)
{
    if (N < 0) return fail(SGP_EINVAL, "N must be >= 0");
    if (d < 1 || d > SGP_MAX_DIM) return fail(SGP_EUNSUPPORTED, "d=%d outside [1, %d]", d, SGP_MAX_DIM);
    if ((double)N * (double)(d + 1) >= 4294967295.0)
        return fail(SGP_EOVERFLOW, "N*(d+1) = %.0f does not fit 32-bit point-vertex ids", (double)N * (d + 1));
    return SGP_OK;
}

extern "C" int sgp_build_points(const float *x, int64_t N, int d, int64_t ldx, const float *scale,
                                int16_t *greedy, int8_t *rank, int32_t *replay, int32_t *status_flags,
                                sgp_stream_t stream)
{
    SGP_RANGE("sgp_build_points");
    int rc = check_dims(N, d);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!x || !scale || !greedy || !rank || !replay || !status_flags || ldx < d)
        return fail(SGP_EINVAL, "sgp_build_points: null pointer or ldx < d");
    ScaleParam sp;
    memset(&sp, 0, sizeof(sp));
    memcpy(sp.s, scale, sizeof(float) * d);
    cudaStream_t st = (cudaStream_t)stream;
    SGP_DISPATCH_D(d, (sgp_points_kernel<DD><<<grid_for(N, 128), 128, 0, st>>>(x, N, d, ldx, sp, greedy, rank, replay,
                                                                               status_flags)));
    return launch_ok("sgp_points_kernel");
}

extern "C" int64_t sgp_hash_capacity(int64_t n_keys)
{
    int64_t cap = 1024;
    while (cap < 2 * n_keys) cap <<= 1;
    return cap;
}

extern "C" int sgp_hash_insert(const int16_t *greedy, const int8_t *rank, int64_t N, int d,
                               uint64_t *table, int64_t capacity, uint32_t *slot_of,
                               int32_t *status_flags, sgp_stream_t stream)
{
    SGP_RANGE("sgp_hash_insert");
    int rc = check_dims(N, d);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!greedy || !rank || !table || !slot_of || !status_flags) return fail(SGP_EINVAL, "sgp_hash_insert: null pointer");
    if (capacity < 2 || (capacity & (capacity - 1)) != 0 || capacity > (1ll << 32))
        return fail(SGP_EINVAL, "hash capacity must be a power of two <= 2^32");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (d + 1);
    SGP_DISPATCH_D(d, (sgp_insert_kernel<DD><<<grid_for(total, 256), 256, 0, st>>>(
                          greedy, rank, N, d, (unsigned long long *)table, (uint64_t)(capacity - 1), slot_of,
                          status_flags)));
    return launch_ok("sgp_insert_kernel");
}

// workspace layout: [marks/pos: total uint32][tile sums: n_tiles uint32][total: uint64 (8-aligned)]
static void number_ws_layout(int64_t total, size_t *off_tiles, size_t *off_total, size_t *bytes)
{
    const int64_t n_tiles = (total + SCAN_TILE - 1) / SCAN_TILE;
    size_t o = (size_t)total * 4;
    *off_tiles = o;
    o += (size_t)(n_tiles > 0 ? n_tiles : 1) * 4;
    o = (o + 15) & ~(size_t)15;
    *off_total = o;
    o += 16;
    *bytes = o;
}

extern "C" size_t sgp_number_workspace_bytes(int64_t N, int d)
{
    size_t a, b, bytes;
    number_ws_layout(N * (int64_t)(d + 1), &a, &b, &bytes);
    return bytes;
}

extern "C" int sgp_count_points(const uint64_t *table, int64_t capacity, const uint32_t *slot_of,
                                int64_t N, int d, void *workspace, size_t workspace_bytes,
                                const int32_t *status_flags, int64_t *M_out, int32_t *flags_out,
                                sgp_stream_t stream)
{
    SGP_RANGE("sgp_count_points");
    int rc = check_dims(N, d);
    if (rc) return rc;
    if (!M_out || !flags_out) return fail(SGP_EINVAL, "sgp_count_points: null output");
    *M_out = 0;
    *flags_out = 0;
    if (N == 0) return SGP_OK;
    (void)capacity;
    const int64_t total = N * (d + 1);
    size_t off_tiles, off_total, need;
    number_ws_layout(total, &off_tiles, &off_total, &need);
    if (!table || !slot_of || !workspace || !status_flags || workspace_bytes < need)
        return fail(SGP_EINVAL, "sgp_count_points: null pointer or workspace too small (%zu < %zu)", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *marks = (uint32_t *)workspace;
    uint32_t *tiles = (uint32_t *)((char *)workspace + off_tiles);
    unsigned long long *total_dev = (unsigned long long *)((char *)workspace + off_total);
    sgp_mark_kernel<<<grid_for(total, 256), 256, 0, st>>>((const unsigned long long *)table, slot_of, total, marks);
    rc = launch_ok("sgp_mark_kernel");
    if (rc) return rc;
    rc = sgp_exclusive_scan_u32(marks, total, tiles, total_dev, st);
    if (rc) return rc;
    unsigned long long m_host = 0;
    int32_t f_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&m_host, total_dev, sizeof(m_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&f_host, status_flags, sizeof(f_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *M_out = (int64_t)m_host;
    *flags_out = f_host;
    if (f_host & SGP_FLAG_TABLE_FULL) return fail(SGP_EOVERFLOW, "hash table full (capacity %lld)", (long long)capacity);
    if (f_host & SGP_FLAG_KEY_RANGE)
        return fail(SGP_ERANGE, "a lattice coordinate does not fit int16: inputs too large for the lengthscale");
    return SGP_OK;
}

extern "C" int sgp_number_points(uint64_t *table, int64_t capacity, const uint32_t *slot_of,
                                 const int16_t *greedy, const int8_t *rank, int64_t N, int d,
                                 const void *workspace, int64_t M, int32_t *replay, int16_t *keys,
                                 sgp_stream_t stream)
{
    SGP_RANGE("sgp_number_points");
    int rc = check_dims(N, d);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!table || !slot_of || !greedy || !rank || !workspace || !replay || !keys)
        return fail(SGP_EINVAL, "sgp_number_points: null pointer");
    if (M < 1 || M > N * (int64_t)(d + 1) || M >= (1ll << 31)) return fail(SGP_EINVAL, "M=%lld out of range", (long long)M);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (d + 1);
    const uint32_t *pos = (const uint32_t *)workspace;
    SGP_DISPATCH_D(d, (sgp_renumber_kernel<DD><<<grid_for(total, 256), 256, 0, st>>>(
                          (const unsigned long long *)table, slot_of, pos, greedy, rank, N, d, replay, keys)));
    rc = launch_ok("sgp_renumber_kernel");
    if (rc) return rc;
    sgp_retarget_kernel<<<grid_for(capacity, 256), 256, 0, st>>>((unsigned long long *)table, capacity, pos);
    return launch_ok("sgp_retarget_kernel");
}

// ---- extension ---------------------------------------------------------------------
static int check_capacity(int64_t capacity)
{
    if (capacity < 2 || (capacity & (capacity - 1)) != 0 || capacity > (1ll << 32))
        return fail(SGP_EINVAL, "hash capacity must be a power of two <= 2^32");
    return SGP_OK;
}

extern "C" int sgp_hash_seed(const int16_t *keys, int64_t M, int d, uint64_t *table, int64_t capacity,
                             int32_t *status_flags, sgp_stream_t stream)
{
    SGP_RANGE("sgp_hash_seed");
    if (d < 1 || d > SGP_MAX_DIM) return fail(SGP_EUNSUPPORTED, "d=%d outside [1, %d]", d, SGP_MAX_DIM);
    if (M < 0 || M >= (1ll << 31)) return fail(SGP_EINVAL, "M=%lld out of range", (long long)M);
    int rc = check_capacity(capacity);
    if (rc) return rc;
    if (M == 0) return SGP_OK;
    if (!keys || !table || !status_flags) return fail(SGP_EINVAL, "sgp_hash_seed: null pointer");
    if (capacity < M) return fail(SGP_EINVAL, "hash capacity %lld below M=%lld", (long long)capacity, (long long)M);
    cudaStream_t st = (cudaStream_t)stream;
    SGP_DISPATCH_D(d, (sgp_seed_kernel<DD><<<grid_for(M, 256), 256, 0, st>>>(keys, M, d, (unsigned long long *)table,
                                                                             (uint64_t)(capacity - 1), status_flags)));
    return launch_ok("sgp_seed_kernel");
}

static int check_extension(int64_t N_new, int d, int64_t M_old)
{
    int rc = check_dims(N_new, d);
    if (rc) return rc;
    if (N_new * (int64_t)(d + 1) >= (1ll << 31))
        return fail(SGP_EOVERFLOW, "N_new*(d+1) = %lld does not fit 31-bit point-vertex ids", (long long)(N_new * (d + 1)));
    if (M_old < 0 || M_old + N_new * (int64_t)(d + 1) >= (1ll << 31))
        return fail(SGP_EOVERFLOW, "M_old + N_new*(d+1) does not fit 31-bit lattice indices");
    return SGP_OK;
}

extern "C" int sgp_hash_extend(const int16_t *greedy_new, const int8_t *rank_new, int64_t N_new, int d,
                               const int16_t *keys, int64_t M_old, uint64_t *table, int64_t capacity,
                               uint32_t *slot_of, int32_t *status_flags, sgp_stream_t stream)
{
    SGP_RANGE("sgp_hash_extend");
    int rc = check_extension(N_new, d, M_old);
    if (rc) return rc;
    rc = check_capacity(capacity);
    if (rc) return rc;
    if (N_new == 0) return SGP_OK;
    if (!greedy_new || !rank_new || !table || !slot_of || !status_flags || (M_old > 0 && !keys))
        return fail(SGP_EINVAL, "sgp_hash_extend: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N_new * (d + 1);
    SGP_DISPATCH_D(d, (sgp_extend_insert_kernel<DD><<<grid_for(total, 256), 256, 0, st>>>(
                          greedy_new, rank_new, N_new, d, keys, (unsigned long long *)table, (uint64_t)(capacity - 1),
                          slot_of, status_flags)));
    return launch_ok("sgp_extend_insert_kernel");
}

extern "C" int sgp_count_extension(const uint64_t *table, int64_t capacity, const uint32_t *slot_of, int64_t N_new,
                                   int d, void *workspace, size_t workspace_bytes, const int32_t *status_flags,
                                   int64_t *M_add_out, int32_t *flags_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_count_extension");
    int rc = check_extension(N_new, d, 0);
    if (rc) return rc;
    if (!M_add_out || !flags_out) return fail(SGP_EINVAL, "sgp_count_extension: null output");
    *M_add_out = 0;
    *flags_out = 0;
    if (N_new == 0) return SGP_OK;
    const int64_t total = N_new * (d + 1);
    size_t off_tiles, off_total, need;
    number_ws_layout(total, &off_tiles, &off_total, &need);
    if (!table || !slot_of || !workspace || !status_flags || workspace_bytes < need)
        return fail(SGP_EINVAL, "sgp_count_extension: null pointer or workspace too small (%zu < %zu)", workspace_bytes,
                    need);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *marks = (uint32_t *)workspace;
    uint32_t *tiles = (uint32_t *)((char *)workspace + off_tiles);
    unsigned long long *total_dev = (unsigned long long *)((char *)workspace + off_total);
    sgp_extend_mark_kernel<<<grid_for(total, 256), 256, 0, st>>>((const unsigned long long *)table, slot_of, total, marks);
    rc = launch_ok("sgp_extend_mark_kernel");
    if (rc) return rc;
    rc = sgp_exclusive_scan_u32(marks, total, tiles, total_dev, st);
    if (rc) return rc;
    unsigned long long m_host = 0;
    int32_t f_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&m_host, total_dev, sizeof(m_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&f_host, status_flags, sizeof(f_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *M_add_out = (int64_t)m_host;
    *flags_out = f_host;
    if (f_host & SGP_FLAG_TABLE_FULL) return fail(SGP_EOVERFLOW, "hash table full (capacity %lld)", (long long)capacity);
    if (f_host & SGP_FLAG_KEY_RANGE)
        return fail(SGP_ERANGE, "a lattice coordinate does not fit int16: inputs too large for the lengthscale");
    return SGP_OK;
}

extern "C" int sgp_number_extension(uint64_t *table, int64_t capacity, const uint32_t *slot_of,
                                    const int16_t *greedy_new, const int8_t *rank_new, int64_t N_new, int d,
                                    const void *workspace, int64_t M_old, int64_t M_add, int32_t *replay_new,
                                    int16_t *keys, sgp_stream_t stream)
{
    SGP_RANGE("sgp_number_extension");
    int rc = check_extension(N_new, d, M_old);
    if (rc) return rc;
    if (N_new == 0) return SGP_OK;
    if (!table || !slot_of || !greedy_new || !rank_new || !workspace || !replay_new || !keys)
        return fail(SGP_EINVAL, "sgp_number_extension: null pointer");
    if (M_add < 0 || M_add > N_new * (int64_t)(d + 1)) return fail(SGP_EINVAL, "M_add=%lld out of range", (long long)M_add);
    (void)capacity;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N_new * (d + 1);
    const uint32_t *pos = (const uint32_t *)workspace;
    SGP_DISPATCH_D(d, (sgp_extend_renumber_kernel<DD><<<grid_for(total, 256), 256, 0, st>>>(
                          (const unsigned long long *)table, slot_of, pos, greedy_new, rank_new, N_new, d,
                          (uint32_t)M_old, replay_new, keys)));
    rc = launch_ok("sgp_extend_renumber_kernel");
    if (rc) return rc;
    sgp_extend_retarget_kernel<<<grid_for(total, 256), 256, 0, st>>>((unsigned long long *)table, slot_of, total, pos,
                                                                     (uint32_t)M_old);
    return launch_ok("sgp_extend_retarget_kernel");
}

// ---- merging key lists (point-sharded build) -------------------------------------------
extern "C" int sgp_hash_append_keys(const int16_t *new_keys, int64_t m_new, int d, const int16_t *keys, int64_t M_old,
                                    uint64_t *table, int64_t capacity, uint32_t *slot_of_new, int32_t *status_flags,
                                    sgp_stream_t stream)
{
    SGP_RANGE("sgp_hash_append_keys");
    if (d < 1 || d > SGP_MAX_DIM) return fail(SGP_EUNSUPPORTED, "d=%d outside [1, %d]", d, SGP_MAX_DIM);
    if (m_new < 0 || M_old < 0 || M_old + m_new >= (1ll << 31))
        return fail(SGP_EOVERFLOW, "M_old + m_new does not fit 31-bit lattice indices");
    int rc = check_capacity(capacity);
    if (rc) return rc;
    if (m_new == 0) return SGP_OK;
    if (!new_keys || !table || !slot_of_new || !status_flags || (M_old > 0 && !keys))
        return fail(SGP_EINVAL, "sgp_hash_append_keys: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    sgp_append_insert_kernel<<<grid_for(m_new, 256), 256, 0, st>>>(new_keys, m_new, d, keys, (unsigned long long *)table,
                                                                   (uint64_t)(capacity - 1), slot_of_new, status_flags);
    return launch_ok("sgp_append_insert_kernel");
}

extern "C" int sgp_count_appended(const uint64_t *table, int64_t capacity, const uint32_t *slot_of_new, int64_t m_new,
                                  void *workspace, size_t workspace_bytes, const int32_t *status_flags,
                                  int64_t *M_add_out, int32_t *flags_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_count_appended");
    if (!M_add_out || !flags_out) return fail(SGP_EINVAL, "sgp_count_appended: null output");
    *M_add_out = 0;
    *flags_out = 0;
    if (m_new == 0) return SGP_OK;
    if (m_new < 0 || m_new >= (1ll << 31)) return fail(SGP_EINVAL, "sgp_count_appended: m_new out of range");
    size_t off_tiles, off_total, need;
    number_ws_layout(m_new, &off_tiles, &off_total, &need);
    if (!table || !slot_of_new || !workspace || !status_flags || workspace_bytes < need)
        return fail(SGP_EINVAL, "sgp_count_appended: null pointer or workspace too small (%zu < %zu)", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *marks = (uint32_t *)workspace;
    uint32_t *tiles = (uint32_t *)((char *)workspace + off_tiles);
    unsigned long long *total_dev = (unsigned long long *)((char *)workspace + off_total);
    // a listed key is new iff its slot is owned by its own list position: the marks of sgp_extend_mark_kernel
    sgp_extend_mark_kernel<<<grid_for(m_new, 256), 256, 0, st>>>((const unsigned long long *)table, slot_of_new, m_new, marks);
    int rc = launch_ok("sgp_extend_mark_kernel");
    if (rc) return rc;
    rc = sgp_exclusive_scan_u32(marks, m_new, tiles, total_dev, st);
    if (rc) return rc;
    unsigned long long m_host = 0;
    int32_t f_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&m_host, total_dev, sizeof(m_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&f_host, status_flags, sizeof(f_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *M_add_out = (int64_t)m_host;
    *flags_out = f_host;
    if (f_host & SGP_FLAG_TABLE_FULL) return fail(SGP_EOVERFLOW, "hash table full (capacity %lld)", (long long)capacity);
    return SGP_OK;
}

extern "C" int sgp_number_appended(uint64_t *table, int64_t capacity, const uint32_t *slot_of_new, const int16_t *new_keys,
                                   int64_t m_new, int d, const void *workspace, int64_t M_old, int64_t M_add,
                                   int32_t *map_out, int16_t *keys, sgp_stream_t stream)
{
    SGP_RANGE("sgp_number_appended");
    if (m_new == 0) return SGP_OK;
    if (!table || !slot_of_new || !new_keys || !workspace || !keys || d < 1 || d > SGP_MAX_DIM || m_new < 0 || M_old < 0 ||
        M_add < 0 || M_add > m_new)
        return fail(SGP_EINVAL, "sgp_number_appended: bad argument");
    (void)capacity;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t *pos = (const uint32_t *)workspace;
    sgp_append_number_kernel<<<grid_for(m_new, 256), 256, 0, st>>>((const unsigned long long *)table, slot_of_new, pos, new_keys,
                                                                   m_new, d, (uint32_t)M_old, map_out, keys);
    int rc = launch_ok("sgp_append_number_kernel");
    if (rc) return rc;
    sgp_extend_retarget_kernel<<<grid_for(m_new, 256), 256, 0, st>>>((unsigned long long *)table, slot_of_new, m_new, pos,
                                                                     (uint32_t)M_old);
    return launch_ok("sgp_extend_retarget_kernel");
}

extern "C" int sgp_build_neighbours(const int16_t *keys, int64_t M, int d, int order,
                                    const uint64_t *table, int64_t capacity, int32_t *nbr,
                                    sgp_stream_t stream)
{
    SGP_RANGE("sgp_build_neighbours");
    if (d < 1 || d > SGP_MAX_DIM) return fail(SGP_EUNSUPPORTED, "d=%d outside [1, %d]", d, SGP_MAX_DIM);
    if (order < 0 || order > SGP_MAX_ORDER) return fail(SGP_EUNSUPPORTED, "order=%d outside [0, %d]", order, SGP_MAX_ORDER);
    if (M == 0 || order == 0) return SGP_OK;
    if (!keys || !table || !nbr || M < 0) return fail(SGP_EINVAL, "sgp_build_neighbours: null pointer");
    if (capacity < 2 || (capacity & (capacity - 1)) != 0) return fail(SGP_EINVAL, "hash capacity must be a power of two");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t work = M * (d + 1);
    SGP_DISPATCH_D(d, (sgp_neighbours_kernel<DD><<<grid_for(work, 256), 256, 0, st>>>(
                          keys, M, d, order, (const unsigned long long *)table, (uint64_t)(capacity - 1), nbr)));
    return launch_ok("sgp_neighbours_kernel");
}

// ---- MVM ---------------------------------------------------------------------------
static int check_view(const sgp_lattice_view *lat, int L)
{
    if (!lat) return fail(SGP_EINVAL, "null lattice view");
    if (lat->replay_stride != 0 && (lat->replay_stride < lat->d + 1 || lat->replay_transposed))
        return fail(SGP_EINVAL, "replay_stride must be 0 or >= d+1 (and the table not transposed)");
    if (lat->N < 0 || lat->M < 0 || lat->d < 1 || lat->d > SGP_MAX_DIM || lat->order < 0 || lat->order > SGP_MAX_ORDER)
        return fail(SGP_EINVAL, "bad lattice view (N=%lld M=%lld d=%d order=%d)", (long long)lat->N, (long long)lat->M,
                    lat->d, lat->order);
    if (L < 1) return fail(SGP_EINVAL, "L must be >= 1");
    if ((double)lat->M * L >= 9.2e18 || (double)lat->N * L >= 9.2e18) return fail(SGP_EOVERFLOW, "M*L overflows");
    return SGP_OK;
}

static inline int pick_vec(int L, int64_t ld_a, int64_t ld_b, const void *p0, const void *p1, const void *p2)
{
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    if (L % 4 == 0 && ld_a % 4 == 0 && ld_b % 4 == 0 && al(p0, 16) && al(p1, 16) && al(p2, 16)) return 4;
    if (L % 2 == 0 && ld_a % 2 == 0 && ld_b % 2 == 0 && al(p0, 8) && al(p1, 8) && al(p2, 8)) return 2;
    return 1;
}

#define SGP_DISPATCH_VEC(vec, CALL)                         \
    switch (vec) {                                          \
        case 4: { constexpr int VV = 4; CALL; } break;      \
        case 2: { constexpr int VV = 2; CALL; } break;      \
        default: { constexpr int VV = 1; CALL; } break;     \
    }

extern "C" int sgp_splat(const sgp_lattice_view *lat, const float *src, int64_t lds, int L,
                         float *values, int mode, sgp_stream_t stream)
{
    SGP_RANGE("sgp_splat");
    int rc = check_view(lat, L);
    if (rc) return rc;
    if (lat->M == 0) return SGP_OK;
    if (!src || !values || !lat->replay || lds < L) return fail(SGP_EINVAL, "sgp_splat: null pointer or lds < L");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SGP_SPLAT_AUTO) mode = (lat->csr_ptr && lat->csr_ent) ? SGP_SPLAT_GATHER : SGP_SPLAT_ATOMIC;
    // accumulate: scatter into `values` as they are (no memset) -- a view of a point range adds its part to a splat that
    // is fed chunk by chunk (sgp_filter_host: the rows of src are splatted while later rows are still being uploaded)
    const bool accumulate = mode == SGP_SPLAT_ATOMIC_ACCUMULATE;
    if (accumulate) mode = SGP_SPLAT_ATOMIC;
    const int vec = pick_vec(L, lds, L, src, values, values);
    const int chunks = L / vec;
    if (mode == SGP_SPLAT_GATHER) {
        if (!lat->csr_ptr || !lat->csr_ent) return fail(SGP_EINVAL, "gather splat needs the CSR arrays");
        const int64_t work = lat->M * chunks;
        SGP_DISPATCH_VEC(vec, (sgp_splat_gather_kernel<VV><<<grid_for(work, 256), 256, 0, st>>>(
                                  lat->csr_ptr, (const int2 *)lat->csr_ent, src, lds, lat->M, L, chunks,
                                  values)));
        return launch_ok("sgp_splat_gather_kernel");
    }
    if (mode != SGP_SPLAT_ATOMIC) return fail(SGP_EINVAL, "unknown splat mode %d", mode);
    if (!accumulate) CUDA_TRY(cudaMemsetAsync(values, 0, sizeof(float) * (size_t)lat->M * (size_t)L, st));
    const int64_t work = lat->N * chunks;
    SGP_DISPATCH_VEC(vec, (sgp_splat_atomic_kernel<VV><<<grid_for(work, 256), 256, 0, st>>>(
                              (const int2 *)lat->replay, lat->replay_transposed ? 1 : (lat->replay_stride > 0 ? lat->replay_stride : lat->d + 1),
                              lat->replay_transposed ? lat->N : 1, lat->perm, src, lds, lat->N, lat->d + 1, L, chunks,
                              values)));
    return launch_ok("sgp_splat_atomic_kernel");
}

extern "C" int sgp_blur(const sgp_lattice_view *lat, const float *coeffs, int k, int L,
                        float *buf0, float *buf1, int *result_in_buf1, sgp_stream_t stream)
{
    SGP_RANGE("sgp_blur");
    int rc = check_view(lat, L);
    if (rc) return rc;
    if (!coeffs || k != 2 * lat->order + 1) return fail(SGP_EINVAL, "stencil length %d does not match order %d", k, lat->order);
    if (result_in_buf1) *result_in_buf1 = 0;
    if (lat->M == 0) return SGP_OK;
    if (!buf0 || !buf1 || (lat->order > 0 && !lat->nbr)) return fail(SGP_EINVAL, "sgp_blur: null pointer");
    (void)rc;
    cudaStream_t st = (cudaStream_t)stream;
    CoeffParam cf;
    memset(&cf, 0, sizeof(cf));
    memcpy(cf.c, coeffs, sizeof(float) * k);
    int vec = pick_vec(L, L, L, buf0, buf1, buf1);
    int chunks = L / vec;
    // a block covers whole rows: block size = largest multiple of `chunks` <= 256 (vec drops to 1 for huge L)
    if (chunks > 256) return fail(SGP_EUNSUPPORTED, "L=%d too wide for one blur block; split the channels", L);
    const int block = (256 / chunks) * chunks;
    const int rows_per_block = block / chunks;
    const int r = lat->order;
    // rows per thread (independent index->row load chains in flight).  Measured on B200 at config A:
    // 1 row/thread 12.6 us per pass, 4 rows/thread 16.9 us -- the pass is L2-bandwidth bound, not latency
    // bound, so the default stays 1.  SGP_BLUR_ROWS=2|4 is a tuning hook for order-1 stencils.
    static int rows_env = 0;
    if (rows_env == 0) {
        const char *e = getenv("SGP_BLUR_ROWS");
        rows_env = e ? atoi(e) : 1;
        if (rows_env != 2 && rows_env != 4) rows_env = 1;
    }
    const int rows = (r == 1) ? rows_env : 1;
    const unsigned grid = (unsigned)((lat->M + (int64_t)rows_per_block * rows - 1) / ((int64_t)rows_per_block * rows));
#define SGP_BLUR_LAUNCH(RR_, ROWS_)                                                                                     \
    do {                                                                                                                \
        if (lat->fast) {                                                                                                \
            SGP_DISPATCH_VEC(vec, (sgp_blur_kernel<VV, RR_, ROWS_, true><<<grid, block, 0, st>>>(nbr_j, in, out, lat->M, L, chunks, r, cf))); \
        } else {                                                                                                        \
            SGP_DISPATCH_VEC(vec, (sgp_blur_kernel<VV, RR_, ROWS_, false><<<grid, block, 0, st>>>(nbr_j, in, out, lat->M, L, chunks, r, cf))); \
        }                                                                                                               \
    } while (0)
    float *in = buf0, *out = buf1;
    for (int j = 0; j <= lat->d; ++j) {
        const int32_t *nbr_j = lat->nbr + (int64_t)j * lat->M * (2 * r);
        if (r == 0) {
            // order-0 stencil: out = c[0] * in
            SGP_BLUR_LAUNCH(0, 1);
        } else if (r == 1 && rows == 4) {
            SGP_BLUR_LAUNCH(1, 4);
        } else if (r == 1 && rows == 2) {
            SGP_BLUR_LAUNCH(1, 2);
        } else if (r == 1) {
            SGP_BLUR_LAUNCH(1, 1);
        } else if (r == 2) {
            SGP_BLUR_LAUNCH(2, 1);
        } else if (r == 3) {
            SGP_BLUR_LAUNCH(3, 1);
        } else {
            SGP_BLUR_LAUNCH(0, 1);
        }
        float *t = in; in = out; out = t;
    }
    if (result_in_buf1) *result_in_buf1 = (in == buf1) ? 1 : 0;
    return launch_ok("sgp_blur_kernel");
}

extern "C" int sgp_slice(const sgp_lattice_view *lat, const float *values, int L, float *out,
                         int64_t ldo, int L_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_slice");
    int rc = check_view(lat, L);
    if (rc) return rc;
    if (lat->N == 0) return SGP_OK;
    if (!values || !out || !lat->replay || L_out < 1 || L_out > L || ldo < L_out)
        return fail(SGP_EINVAL, "sgp_slice: null pointer, L_out outside [1, L] or ldo < L_out");
    // production form: replay table through warp-private TMA rings (sgp_ring.cu); SGP_RING=0 selects the one-shot kernel
    if (sgp_ring_slice_enabled() && sgp_slice_ring_supported(lat, values, L) && !(L_out != L && L % 2 != 0))
        return sgp_slice_ring(lat, values, L, out, ldo, L_out, stream);
    cudaStream_t st = (cudaStream_t)stream;
    // the lattice side decides the vector width; out is written channel by channel when it does not match it
    int vec = pick_vec(L, L, L, values, values, values);
    const bool ragged = vec > 1 && !(L_out == L && ldo % vec == 0 && ((uintptr_t)out % (4 * vec)) == 0);
    if (vec == 1 && L_out != L) return fail(SGP_EUNSUPPORTED, "sgp_slice: L_out < L needs vectorisable lattice rows");
    const int chunks = L / vec;
    const int64_t work = lat->N * chunks;
    const float divisor = sgp_slice_divisor(lat->d);
    volatile float rdivisor = 1.0f / divisor;
    static int batch = 0;   // tuning hook: SGP_SLICE_BATCH=3|9 (vertices whose loads are in flight together)
    if (batch == 0) {
        const char *e = getenv("SGP_SLICE_BATCH");
        batch = (e && atoi(e) == 3) ? 3 : 9;
    }
#define SGP_SLICE_LAUNCH(BB, FF, SS, RG)                                                                                     \
    SGP_DISPATCH_VEC(vec, (launch_err = sgp_launch_pdl(sgp_slice_kernel<VV, BB, FF, SS, RG>, dim3(grid_for(work, 256)), dim3(256), 0, st, \
                              (const int2 *)lat->replay, (int64_t)(lat->replay_transposed ? 1 : (lat->replay_stride > 0 ? lat->replay_stride : lat->d + 1)), \
                              (int64_t)(lat->replay_transposed ? lat->N : 1), lat->perm, values, lat->N, lat->d + 1, L, chunks, \
                              divisor, (float)rdivisor, out, ldo, L_out)))
    cudaError_t launch_err = cudaSuccess;
    static int stream_env = -1;   // SGP_SLICE_STREAM=0 turns off the streaming cache policy of the replay reads / out writes (69 -> 66 us with it)
    if (stream_env < 0) {
        const char *e = getenv("SGP_SLICE_STREAM");
        stream_env = e ? atoi(e) : 1;
    }
    if (ragged) {
        if (lat->fast) { SGP_SLICE_LAUNCH(9, true, true, true); } else { SGP_SLICE_LAUNCH(9, false, true, true); }
    } else if (batch == 3) {
        if (lat->fast) { SGP_SLICE_LAUNCH(3, true, false, false); } else { SGP_SLICE_LAUNCH(3, false, false, false); }
    } else if (stream_env) {
        if (lat->fast) { SGP_SLICE_LAUNCH(9, true, true, false); } else { SGP_SLICE_LAUNCH(9, false, true, false); }
    } else {
        if (lat->fast) { SGP_SLICE_LAUNCH(9, true, false, false); } else { SGP_SLICE_LAUNCH(9, false, false, false); }
    }
#undef SGP_SLICE_LAUNCH
    if (launch_err != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_slice_kernel failed: %s", cudaGetErrorString(launch_err));
    return launch_ok("sgp_slice_kernel");
}

// Test hook: counts a in [lo, lo+count) (as fp32 bit patterns) for which the Markstein division used by
// slice differs from the IEEE division by the slice divisor of dimension d.
extern "C" int sgp_debug_division_mismatches(int d, uint32_t lo, uint32_t count, unsigned long long *mismatches_dev,
                                             sgp_stream_t stream)
{
    if (!mismatches_dev || count == 0) return fail(SGP_EINVAL, "sgp_debug_division_mismatches: bad argument");
    const float divisor = sgp_slice_divisor(d);
    volatile float rdivisor = 1.0f / divisor;
    sgp_exact_div_check_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(divisor, rdivisor, lo, count,
                                                                                         mismatches_dev);
    return launch_ok("sgp_exact_div_check_kernel");
}

extern "C" int sgp_mvm(const sgp_lattice_view *lat, const float *src, int64_t lds, int L,
                       const float *coeffs, int k, float *out, int64_t ldo,
                       float *buf0, float *buf1, int splat_mode, sgp_stream_t stream)
{
    int rc = sgp_splat(lat, src, lds, L, buf0, splat_mode, stream);
    if (rc) return rc;
    int in1 = 0;
    rc = sgp_blur(lat, coeffs, k, L, buf0, buf1, &in1, stream);
    if (rc) return rc;
    return sgp_slice(lat, in1 ? buf1 : buf0, L, out, ldo, L, stream);
}
