// Round-2 micro-benchmark behind DESIGN.md section 3.7: can the 64-byte row gathers of splat / slice go through the TMA
// (one cp.async.bulk per row into shared memory, mbarrier completion) instead of the L1 data pipe?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o profiles/micro/bulk_gather profiles/micro/bulk_gather.cu
//   profiles/micro/bulk_gather [rows=400000] [gathers=9000000]
//
// Both kernels gather `gathers` random 64-byte rows of a [rows, 16] fp32 table (the metric lattice: 398 902 rows, 9M
// gathers per splat or slice) and reduce them to a checksum:
//   ldg  : 4 lanes x ld.global.v4 per row, 8 rows per warp instruction, 4 instructions in flight per thread
//   bulk : persistent warps, 2 stages of 32 rows; each lane issues one 64-byte cp.async.bulk for its row, the warp waits on
//          the stage's mbarrier and reads the rows back from shared memory (half the L1-pipe wavefronts of the ldg form)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) gather_ldg(const int *__restrict__ idx, const float4 *__restrict__ table, int64_t n,
                                                  float *__restrict__ out)
{
    const int lane4 = threadIdx.x & 3;
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 2;
    float acc = 0.0f;
    for (; g + 3 * stride < n; g += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = table[(int64_t)idx[g + k * stride] * 4 + lane4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(256) gather_bulk(const int *__restrict__ idx, const float *__restrict__ table, int64_t n,
                                                   float *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *rows = (float *)smem + (size_t)warp * STAGES * 32 * 16;                  // [STAGES][32 rows][16 floats]
    uint64_t *bars = (uint64_t *)(smem + (size_t)(blockDim.x >> 5) * STAGES * 2048) + warp * STAGES;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t batches = n / 32;
    auto issue = [&](int64_t b, int s) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(2048u) : "memory");
        __syncwarp();
        const int r = idx[b * 32 + lane];
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                     ::"r"(smem_u32(rows + (s * 32 + lane) * 16)), "l"(table + (int64_t)r * 16), "r"(smem_u32(bars + s)) : "memory");
    };
    int64_t b = w0;
    for (int s = 0; s < STAGES && b + s * warps < batches; ++s) issue(b + s * warps, s);
    float acc = 0.0f;
    uint32_t phase = 0;
    int s = 0;
    for (; b < batches; b += warps) {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bars + s)), "r"(phase) : "memory");
        } while (!done);
        const float4 *p = (const float4 *)(rows + s * 32 * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 v = p[k * 32 + lane]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
        if (b + STAGES * warps < batches) issue(b + STAGES * warps, s);
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char **argv)
{
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 400000, n = (argc > 2 ? atoll(argv[2]) : 9000000) / 32 * 32;
    std::vector<int> h(n);
    uint64_t sd = 88172645463325252ull;
    for (auto &v : h) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; v = (int)(sd % (uint64_t)rows); }
    int *idx; float *table, *out;
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&table, rows * 64)); CK(cudaMalloc(&out, 4));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(table, 0, rows * 64));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto report = [&](const char *name, float ms, int reps) {
        const double us = ms / reps * 1e3;
        printf("%-22s %8.1f us per %lld gathers  = %6.2f G rows/s, %5.2f TB/s of rows\n", name, us, (long long)n, n / us * 1e-3, n * 64 / us * 1e-6);
    };
    const int reps = 20;
    float ms;
    for (int i = 0; i < 3; ++i) gather_ldg<<<sms * 8, 256>>>(idx, (const float4 *)table, n, out);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) gather_ldg<<<sms * 8, 256>>>(idx, (const float4 *)table, n, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    report("ld.global.v4 (8 CTA/SM)", ms, reps);
#define RUN_BULK(ST, CTAS)                                                                                              \
    {                                                                                                                   \
        const size_t sm = 8 * ST * 2048 + 8 * ST * 8;                                                                   \
        CK(cudaFuncSetAttribute(gather_bulk<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));               \
        for (int i = 0; i < 2; ++i) gather_bulk<ST><<<sms * CTAS, 256, sm>>>(idx, table, n, out);                       \
        CK(cudaEventRecord(e0));                                                                                        \
        for (int i = 0; i < reps; ++i) gather_bulk<ST><<<sms * CTAS, 256, sm>>>(idx, table, n, out);                    \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));                  \
        report("cp.async.bulk " #ST " st x " #CTAS " CTA", ms, reps);                                                    \
    }
    RUN_BULK(2, 2) RUN_BULK(2, 4) RUN_BULK(4, 2)
    CK(cudaDeviceSynchronize());
    return 0;
}
