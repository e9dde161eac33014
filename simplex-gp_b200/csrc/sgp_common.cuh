// sgp_common.cuh -- pieces shared by the translation units of libsgp_lattice.so
#ifndef SGP_COMMON_CUH
#define SGP_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "sgp_lattice.h"

int sgp_fail(int code, const char *fmt, ...);
int sgp_pdl_enabled(void);   // SGP_PDL=0 turns programmatic dependent launch off (default on)
int sgp_launch_ok(const char *what);

// NVTX range per stage (the replacement for the reference's -DDEBUG stage timers, permutohedral.h:268-336): every stage
// entry point opens a range named after itself when the environment says SGP_NVTX=1 (read once); otherwise one branch.
int sgp_nvtx_enabled(void);
void sgp_nvtx_push(const char *name);
void sgp_nvtx_pop(void);
struct SgpRange {
    bool on;
    explicit SgpRange(const char *name) : on(sgp_nvtx_enabled() != 0) { if (on) sgp_nvtx_push(name); }
    ~SgpRange() { if (on) sgp_nvtx_pop(); }
};
#define SGP_RANGE(name) SgpRange sgp_range_guard_(name)

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return sgp_fail(SGP_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));       \
    } while (0)

// exclusive scan of data[n] (uint32) in place; tile_sums: scratch of sgp_scan_tiles(n) uint32;
// grand total -> *total_dev (device uint64).  Three launches on `st`.
#define SGP_SCAN_TILE 4096
static inline int64_t sgp_scan_tiles(int64_t n) { return n > 0 ? (n + SGP_SCAN_TILE - 1) / SGP_SCAN_TILE : 1; }
int sgp_exclusive_scan_u32(uint32_t *data, int64_t n, uint32_t *tile_sums, unsigned long long *total_dev,
                           cudaStream_t st);

int sgp_splat_rows_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                             const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream);
int sgp_cg_reduce_partials(const float *partial, int blocks, int L, float *out, sgp_stream_t stream);
int sgp_splat_rows_ring_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                  const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream);
cudaStream_t sgp_side_stream(int dev);   // one internal non-blocking stream per device (created on first use)

static inline unsigned sgp_grid_for(int64_t work, int block) { return (unsigned)((work + block - 1) / block); }

// Loads whose ISSUE ORDER matters (a batch of index loads, then a batch of dependent row loads, so that a thread has
// several independent memory round trips in flight): volatile asm statements keep their relative order, whereas
// plain __ldg calls get interleaved with their consumers by the compiler.
__device__ __forceinline__ int2 ldg_ordered_int2(const int2 *p)
{
    int2 v;
    asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// same, for data that is read exactly once (replay tables): streaming policy, so that it does not displace the
// lattice values, which the gathers re-read from L2
__device__ __forceinline__ int2 ldg_ordered_int2_streaming(const int2 *p)
{
    int2 v;
    asm volatile("ld.global.cs.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_ordered_f4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_ordered_f2(const float *p)
{
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_ordered_f1(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- layout of the row-sorted entries ---------------------------------------------------------------------------
// The entries {point | row-start flag, weight} are consumed 8 at a time (a "segment", one thread), as four 16-byte pieces.
// They are stored in groups of 8 segments (64 entries, 512 bytes) with the pieces interleaved: [piece 0..3][segment 0..7],
// so that the threads of a warp that read piece i of consecutive segments read ONE contiguous 128-byte run (linear
// storage puts them 64 bytes apart: four 128-byte lines per instruction, and a 4-way bank conflict when staged in
// shared memory).  sgp_entry_index(e) = position (in entries) of row-sorted entry e.
__host__ __device__ __forceinline__ int64_t sgp_entry_index(int64_t e)
{
    return ((e >> 6) << 6) + (((e & 7) >> 1) << 4) + (((e >> 3) & 7) << 1) + (e & 1);
}
#define SGP_ENTRY_GROUP 64   /* entries per interleave group; the arrays are padded to a multiple of it */

// ---- key hash table (shared by the lattice build and the blur-group build) ------------------------------
// A slot is {fingerprint:32 | value:32}; after sgp_number_points the value is the lattice index of the key.
#define SGP_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ uint64_t mix64(uint64_t h)
{
    h ^= h >> 33;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    h *= 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 33;
    return h;
}

// hash of a d-vector of int16 (table layout and hash are not observable: permutohedral.h:114-121
// only has to be *a* hash).  Two coordinates per round.
template <int D, typename KeyArr>
__device__ __forceinline__ uint64_t hash_key(const KeyArr &key, int d)
{
    uint64_t h = 0x9E3779B97F4A7C15ull;
    if (D > 0) {
#pragma unroll
        for (int i = 0; i + 1 < (D > 0 ? D : 1); i += 2) {
            uint32_t w = (uint32_t)(uint16_t)key[i] | ((uint32_t)(uint16_t)key[i + 1] << 16);
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            h ^= h >> 29;
        }
        if (D & 1) {
            uint32_t w = (uint32_t)(uint16_t)key[(D > 0 ? D : 1) - 1];
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            h ^= h >> 29;
        }
    } else {
        int i = 0;
        for (; i + 1 < d; i += 2) {
            uint32_t w = (uint32_t)(uint16_t)key[i] | ((uint32_t)(uint16_t)key[i + 1] << 16);
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            h ^= h >> 29;
        }
        if (i < d) {
            uint32_t w = (uint32_t)(uint16_t)key[i];
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            h ^= h >> 29;
        }
    }
    return mix64(h);
}


// find the lattice index of key nk[0..d) in a table built by sgp_hash_insert + sgp_number_points; -1 if absent
__device__ __forceinline__ int32_t sgp_table_find(const int16_t *__restrict__ keys,
                                                  const unsigned long long *__restrict__ table, uint64_t mask,
                                                  const int16_t *nk, int d)
{
    const uint64_t h = hash_key<0>(nk, d);
    const uint32_t fp = (uint32_t)(h >> 32);
    uint64_t slot = h & mask;
    for (uint64_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long cur = table[slot];
        if (cur == SGP_EMPTY) return -1;
        if ((uint32_t)(cur >> 32) == fp) {
            const uint32_t idx = (uint32_t)cur;
            const int16_t *op = keys + (int64_t)idx * d;
            bool same = true;
            for (int c = 0; c < d; ++c) same = same && (op[c] == nk[c]);
            if (same) return (int32_t)idx;
        }
        slot = (slot + 1) & mask;
    }
    return -1;
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// The MVM is a chain of short kernels (splat -> blur stages -> slice, 20-90 us each).  Each of them starts with work
// that does not depend on its predecessor's output (index-table loads, shared-memory staging of neighbour tables),
// so they are launched with the programmatic-stream-serialization attribute: a kernel calls pdl_launch_dependents()
// at once (its successor's CTAs may then be scheduled as SMs drain) and pdl_wait() just before it first touches data
// its predecessor wrote (that returns only when the predecessor has completed and flushed).  Launched without the
// attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t sgp_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = sgp_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- vector of VEC channels of one row ------------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void load_cg(const float *p) { v[0] = __ldcg(p); }
    __device__ __forceinline__ void load_ordered(const float *p) { v[0] = ldg_ordered_f1(p); }
    __device__ __forceinline__ void load_plain(const float *p) { v[0] = *p; }
    __device__ __forceinline__ void store(float *p) const { *p = v[0]; }
    __device__ __forceinline__ void store_streaming(float *p) const { __stcs(p, v[0]); }
    __device__ __forceinline__ void red(float *p) const { atomicAdd(p, v[0]); }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float *p) { float2 t = __ldg((const float2 *)p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_cg(const float *p) { float2 t = __ldcg((const float2 *)p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_ordered(const float *p) { float2 t = ldg_ordered_f2(p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_plain(const float *p) { float2 t = *(const float2 *)p; v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void store(float *p) const { *(float2 *)p = make_float2(v[0], v[1]); }
    __device__ __forceinline__ void store_streaming(float *p) const { __stcs((float2 *)p, make_float2(v[0], v[1])); }
    __device__ __forceinline__ void red(float *p) const
    {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    }
};
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) { float4 t = __ldg((const float4 *)p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_cg(const float *p) { float4 t = __ldcg((const float4 *)p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_ordered(const float *p) { float4 t = ldg_ordered_f4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_plain(const float *p) { float4 t = *(const float4 *)p; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void store(float *p) const { *(float4 *)p = make_float4(v[0], v[1], v[2], v[3]); }
    __device__ __forceinline__ void store_streaming(float *p) const { __stcs((float4 *)p, make_float4(v[0], v[1], v[2], v[3])); }
    __device__ __forceinline__ void red(float *p) const
    {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
};


// acc + a*b: the reference's two roundings (EXACT) or one fused multiply-add (FAST; differs by rounding only, far
// inside the 1e-5 relative tolerance of the value path)
template <bool FAST> __device__ __forceinline__ float madd(float a, float b, float acc)
{
    return FAST ? __fmaf_rn(a, b, acc) : __fadd_rn(acc, __fmul_rn(a, b));
}

// ---- asynchronous global -> shared row gathers (LDGSTS) --------------------------------------------
// A gather "index -> row" is two dependent memory round trips.  Warps issue in order, so a loop of such pairs
// serialises them; instead the index list is staged in shared memory first and the rows are then fetched with
// cp.async, which costs no registers and never blocks the issuing warp: every row of a CTA's batch is in flight at once.
template <int VEC> __device__ __forceinline__ void cp_async_vec(float *smem_dst, const float *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (VEC == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if (VEC == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// Cooperative asynchronous copy of nbytes (a multiple of 4; both pointers 4-byte aligned, the shared one 16-byte
// aligned) by the whole CTA: 16-byte pieces when the global pointer allows it, 4-byte pieces otherwise.
// Never "smem[i] = global[i]" in a loop: each iteration would stall the warp for a full memory round trip.
__device__ __forceinline__ void cta_copy_async(void *smem_dst, const void *gsrc, int nbytes, int tid, int nthreads)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const char *g = (const char *)gsrc;
    int done = 0;
    if ((((uintptr_t)g) & 15) == 0) {
        const int n16 = nbytes >> 4;
        for (int i = tid; i < n16; i += nthreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * i), "l"(g + 16 * (size_t)i) : "memory");
        done = n16 << 4;
    }
    const int n4 = (nbytes - done) >> 2;
    for (int i = tid; i < n4; i += nthreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + done + 4u * i), "l"(g + done + 4 * (size_t)i) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// a / b for a fixed divisor b whose reciprocal rb = RN(1/b) was computed on the host (Markstein:
// q0 = RN(a*rb), rem = a - q0*b exactly by FMA, q = RN(q0 + rem*rb)).  Equal to the IEEE division
// bit for bit for every finite |a| >= 2^-100 and for a = 0 (sign of zero aside, which cannot reach
// the sum); below 2^-100 the remainder may be inexact and q can be off by one denormal-range ulp
// (absolute error < 1e-37).  tests/test_gpu_parity.py::test_exact_division checks both claims over
// all 2^32 bit patterns.  Five issue slots per term instead of the ~12 of __fdiv_rn.
__device__ __forceinline__ float exact_div(float a, float b, float rb)
{
    const float q0 = __fmul_rn(a, rb);
    const float rem = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(rem, rb, q0);
}


#endif  // SGP_COMMON_CUH
