// sgp_common.cuh -- pieces shared by the translation units of libsgp_lattice.so
#ifndef SGP_COMMON_CUH
#define SGP_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "sgp_lattice.h"

int sgp_fail(int code, const char *fmt, ...);
int sgp_launch_ok(const char *what);

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

static inline unsigned sgp_grid_for(int64_t work, int block) { return (unsigned)((work + block - 1) / block); }

// ---- vector of VEC channels of one row ------------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void load_cg(const float *p) { v[0] = __ldcg(p); }
    __device__ __forceinline__ void load_plain(const float *p) { v[0] = *p; }
    __device__ __forceinline__ void store(float *p) const { *p = v[0]; }
    __device__ __forceinline__ void red(float *p) const { atomicAdd(p, v[0]); }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float *p) { float2 t = __ldg((const float2 *)p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_cg(const float *p) { float2 t = __ldcg((const float2 *)p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void load_plain(const float *p) { float2 t = *(const float2 *)p; v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void store(float *p) const { *(float2 *)p = make_float2(v[0], v[1]); }
    __device__ __forceinline__ void red(float *p) const
    {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    }
};
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) { float4 t = __ldg((const float4 *)p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_cg(const float *p) { float4 t = __ldcg((const float4 *)p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load_plain(const float *p) { float4 t = *(const float4 *)p; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void store(float *p) const { *(float4 *)p = make_float4(v[0], v[1], v[2], v[3]); }
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
