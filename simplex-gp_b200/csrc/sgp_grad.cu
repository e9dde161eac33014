// sgp_grad.cu -- the two element-wise ends of the lengthscale-gradient pass (B200, sm_100a).
//
// Reference: LatticeFilterGeneral.backward, gpytorch_lattice_kernel/bilateral_kernel.py:97-124.  With
// g = grad_output [N, L], v = source [N, L], x = reference [N, d] it filters ONE block
//     all_ = [ g | g (x) x | v | v (x) x ]        N x 2L(1+d) channels          (:113-119)
// with the derivative stencil and contracts
//     grad_x[n, k] = -2 * sum_l ( (v x)[n,l,k] * wg[n,l] - v[n,l] * wgx[n,l,k]
//                               + (g x)[n,l,k] * wv[n,l] - g[n,l] * wvx[n,l,k] )   (:122)
// The lattice filter is linear and acts on every channel independently, so the channels can be
// filtered in any grouping.  Here they are grouped per RHS column l,
//     block(l) = [ g_l, g_l x_1 .. g_l x_d, v_l, v_l x_1 .. v_l x_d ]      2(1+d) channels,
// a "pack" kernel writes the blocks of columns [l0, l0+nl) into one [N, ldp] matrix, the ordinary
// MVM kernels filter it, and a "contract" kernel folds the filtered block into grad_x, visiting
// the columns in increasing l with the reference's left-to-right expression so that the fp32
// result does not depend on how the columns were chunked.  The N x 2L(1+d) matrix of the
// reference never exists; device memory is bounded by the chunk (nl columns).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok
#define grid_for sgp_grid_for

// thread = (point n, column j of the chunk, slot s in [0, 2(1+d)))
__global__ void __launch_bounds__(256)
sgp_grad_pack_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ v, int64_t ldv,
                     const float *__restrict__ x, int64_t ldx, int64_t N, int d, int l0, int nl,
                     float *__restrict__ packed, int64_t ldp)
{
    const int per = 2 * (d + 1);
    const int width = per * nl;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / width;
    if (n >= N) return;
    const int c = (int)(tid - n * width);
    const int j = c / per;
    const int s = c - j * per;
    const bool second = s > d;              // the v half of the block
    const int k = second ? s - (d + 1) : s; // 0 = plain column, 1..d = column times x_k
    const float a = second ? __ldg(v + n * ldv + l0 + j) : __ldg(g + n * ldg + l0 + j);
    const float val = (k == 0) ? a : __fmul_rn(a, __ldg(x + n * ldx + (k - 1)));
    packed[n * ldp + c] = val;
}

// The same block written 16 bytes at a time: thread = (point n, chunk q of four channels); one integer division per
// thread instead of two per channel, one float4 store instead of four scalar ones (0.63 -> 0.1x ms per 144-channel
// block at N = 1M).  Needs ldp % 4 == 0 and a 16-byte aligned block; channels past the chunk's width are written as 0.
__global__ void __launch_bounds__(256)
sgp_grad_pack4_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ v, int64_t ldv,
                      const float *__restrict__ x, int64_t ldx, int64_t N, int d, int l0, int nl,
                      float *__restrict__ packed, int64_t ldp)
{
    const int per = 2 * (d + 1);
    const int width = per * nl;
    const int chunks = (width + 3) >> 2;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / chunks;
    if (n >= N) return;
    const int c0 = (int)(tid - n * chunks) * 4;
    int j = c0 / per;
    int s = c0 - j * per;
    float out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float val = 0.0f;
        if (c0 + i < width) {
            const bool second = s > d;
            const int k = second ? s - (d + 1) : s;
            const float a = second ? __ldg(v + n * ldv + l0 + j) : __ldg(g + n * ldg + l0 + j);
            val = (k == 0) ? a : __fmul_rn(a, __ldg(x + n * ldx + (k - 1)));
        }
        out[i] = val;
        if (++s == per) { s = 0; ++j; }
    }
    __stcs((float4 *)(packed + n * ldp + c0), make_float4(out[0], out[1], out[2], out[3]));
}

// thread = (point n, axis k).  acc[n, k] carries the running sum over columns between chunks.
__global__ void __launch_bounds__(256)
sgp_grad_contract_kernel(const float *__restrict__ filtered, int64_t ldp, const float *__restrict__ g, int64_t ldg,
                         const float *__restrict__ v, int64_t ldv, const float *__restrict__ x, int64_t ldx,
                         int64_t N, int d, int l0, int nl, int first, int last, float *__restrict__ grad_x,
                         int64_t ldgx)
{
    const int per = 2 * (d + 1);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / d;
    if (n >= N) return;
    const int k = (int)(tid - n * d);
    const float xk = __ldg(x + n * ldx + k);
    float acc = first ? 0.0f : grad_x[n * ldgx + k];
    const float *f = filtered + n * ldp;
    for (int j = 0; j < nl; ++j) {
        const float gl = __ldg(g + n * ldg + l0 + j);
        const float vl = __ldg(v + n * ldv + l0 + j);
        const float *b = f + j * per;
        const float wg = b[0], wgx = b[1 + k], wv = b[d + 1], wvx = b[d + 2 + k];
        // ((sf*wg - src*wgf) + gf*ws) - g*wsf, each product rounded separately (no contraction)
        float t = __fsub_rn(__fmul_rn(__fmul_rn(vl, xk), wg), __fmul_rn(vl, wgx));
        t = __fadd_rn(t, __fmul_rn(__fmul_rn(gl, xk), wv));
        t = __fsub_rn(t, __fmul_rn(gl, wvx));
        acc = __fadd_rn(acc, t);
    }
    grad_x[n * ldgx + k] = last ? __fmul_rn(-2.0f, acc) : acc;
}

// column 0 of every block (the filtered g_l) -> grad_source[:, l0:l0+nl]
__global__ void __launch_bounds__(256)
sgp_grad_take_wg_kernel(const float *__restrict__ filtered, int64_t ldp, int64_t N, int d, int l0, int nl,
                        float *__restrict__ grad_src, int64_t ldgs)
{
    const int per = 2 * (d + 1);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / nl;
    if (n >= N) return;
    const int j = (int)(tid - n * nl);
    grad_src[n * ldgs + l0 + j] = filtered[n * ldp + (int64_t)j * per];
}

static int check_common(int64_t N, int d, int l0, int nl, int64_t ldp)
{
    if (N < 0 || d < 1 || d > SGP_MAX_DIM || l0 < 0 || nl < 1) return fail(SGP_EINVAL, "sgp_grad: bad shape");
    if (ldp < (int64_t)2 * (d + 1) * nl) return fail(SGP_EINVAL, "sgp_grad: ldp smaller than 2(1+d)*nl");
    return SGP_OK;
}

extern "C" int sgp_grad_channels(int d, int nl) { return 2 * (d + 1) * nl; }

extern "C" int sgp_grad_pack(const float *g, int64_t ldg, const float *v, int64_t ldv, const float *x, int64_t ldx,
                             int64_t N, int d, int l0, int nl, float *packed, int64_t ldp, sgp_stream_t stream)
{
    SGP_RANGE("sgp_grad_pack");
    int rc = check_common(N, d, l0, nl, ldp);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!g || !v || !x || !packed) return fail(SGP_EINVAL, "sgp_grad_pack: null pointer");
    const int width = 2 * (d + 1) * nl;
    if (ldp % 4 == 0 && ((uintptr_t)packed % 16) == 0 && ldp >= (int64_t)((width + 3) / 4 * 4)) {
        const int64_t work = N * (int64_t)((width + 3) / 4);
        sgp_grad_pack4_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, v, ldv, x, ldx, N, d, l0,
                                                                                   nl, packed, ldp);
        return launch_ok("sgp_grad_pack4_kernel");
    }
    const int64_t work = N * (int64_t)width;
    sgp_grad_pack_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, v, ldv, x, ldx, N, d, l0, nl,
                                                                              packed, ldp);
    return launch_ok("sgp_grad_pack_kernel");
}

extern "C" int sgp_grad_contract(const float *filtered, int64_t ldp, const float *g, int64_t ldg, const float *v,
                                 int64_t ldv, const float *x, int64_t ldx, int64_t N, int d, int l0, int nl,
                                 int first, int last, float *grad_x, int64_t ldgx, float *grad_src, int64_t ldgs,
                                 sgp_stream_t stream)
{
    SGP_RANGE("sgp_grad_contract");
    int rc = check_common(N, d, l0, nl, ldp);
    if (rc) return rc;
    if (N == 0) return SGP_OK;
    if (!filtered || !g || !v || !x || !grad_x || ldgx < d) return fail(SGP_EINVAL, "sgp_grad_contract: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    sgp_grad_contract_kernel<<<grid_for(N * (int64_t)d, 256), 256, 0, st>>>(filtered, ldp, g, ldg, v, ldv, x, ldx, N, d,
                                                                          l0, nl, first, last, grad_x, ldgx);
    rc = launch_ok("sgp_grad_contract_kernel");
    if (rc) return rc;
    if (grad_src) {
        sgp_grad_take_wg_kernel<<<grid_for(N * (int64_t)nl, 256), 256, 0, st>>>(filtered, ldp, N, d, l0, nl, grad_src,
                                                                              ldgs);
        rc = launch_ok("sgp_grad_take_wg_kernel");
    }
    return rc;
}
