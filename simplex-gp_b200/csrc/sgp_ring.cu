// sgp_ring.cu -- splat and slice with their index streams prefetched through warp-private TMA rings (B200, sm_100a).
//
// Reference semantics: PermutohedralLattice::splat accumulation (gpytorch_lattice_kernel/cpp/permutohedral.h:478-479)
// and ::slice (:497-510), on the tables of a built lattice.
//
// Why.  Both kernels are two dependent memory round trips per warp: the index stream (row-sorted entries / replay
// table: contiguous, read once, comes from HBM: ~1000 cycles) and then the 64-byte row gathers it addresses (L2:
// ~400 cycles).  The one-shot kernels (sgp_splat_rows_kernel, sgp_slice_kernel) pay both latencies in every warp and
// the index loads cost a third of their L1 wavefronts (ncu, profiles/r1_mvm_full.txt: 89 k wavefronts per SM, nothing
// above 75 % busy, 13-14 long-scoreboard stalls per issue).  Here a warp is persistent and owns a small ring of
// shared-memory stages; one lane streams the next tiles of the index stream into it with cp.async.bulk (TMA, one
// instruction per tile, completion on an mbarrier, L2 evict-first policy), so by the time the warp turns to a tile
// its indices sit in shared memory: the only exposed latency is the row gather, and the index stream costs no LSU
// wavefronts on the global path.  No CTA-wide barrier anywhere: warps never wait for each other.
//
// The splat (production form for dense lattices at 8..64 columns, see sgp_tiles.cu::splat_rows_impl): a tile is 256
// row-sorted entries in the interleaved storage order (sgp_entry_index: the eight lane groups of a warp read eight
// consecutive 16-byte words) plus its 32 segment rows, delivered by ONE bulk copy pair; every thread reduces its runs
// of equal lattice row into the (pre-zeroed) values with red.global.add.v4.f32, except that a pass whose segments all
// lie inside one lattice row -- the long rows at the centre of the data -- is first summed across the warp with a
// butterfly (segments sit at power-of-two lane strides; with 3, 5, 6, 7 chunks per row the spare lanes idle).
// Optional form (SGP_SPLAT_SCAN=1, measured slower): runs are combined across the threads of a tile with a segmented
// scan and stored; only rows that cross a tile boundary are reduced (and only those are zeroed beforehand,
// sgp_ring_zero_heads_kernel), so the memset of the lattice values is gone -- at the price of 25 shuffles per pass.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "sgp_common.cuh"
#include "sgp_lattice.h"

#define fail sgp_fail
#define launch_ok sgp_launch_ok

#define RING_MAX_STAGES 4
#define RING_THREADS 256
#define RING_WARPS (RING_THREADS / 32)
#define ROW_START_FLAG 0x80000000u

// ---- mbarrier / bulk-copy primitives --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy (TMA, non-tensor form): bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

template <int VEC> __device__ __forceinline__ void vec_zero(Vec<VEC> &a)
{
#pragma unroll
    for (int k = 0; k < VEC; ++k) a.v[k] = 0.0f;
}

// =====================================================================================================
// slice
// =====================================================================================================
// B lattice rows of one point: all B entries are read from shared memory first, then all B rows are requested, then
// the sum runs in vertex order (the reference's order).  The branch on the sign of the indices (never negative) keeps
// the compiler from interleaving the shared-memory reads with the row loads (warps issue in order).
template <int VEC, bool FAST, int B>
__device__ __forceinline__ void slice_batch(const int2 *ep, const float *__restrict__ values, int L, int c0,
                                            Vec<VEC> &acc, float divisor, float rdivisor)
{
    int2 e[B];
    Vec<VEC> v[B];
#pragma unroll
    for (int b = 0; b < B; ++b) e[b] = ep[b];
    int lowest = e[0].x;
#pragma unroll
    for (int b = 1; b < B; ++b) lowest = min(lowest, e[b].x);
    if (lowest >= 0) {
#pragma unroll
        for (int b = 0; b < B; ++b) v[b].load_ordered(values + (int64_t)e[b].x * L + c0);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const float w = __int_as_float(e[b].y);
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                acc.v[k] = FAST ? __fmaf_rn(w, v[b].v[k], acc.v[k])
                                : __fadd_rn(acc.v[k], exact_div(__fmul_rn(w, v[b].v[k]), divisor, rdivisor));
        }
    }
}

// The same with the entries read in PAIRS (16-byte shared-memory loads): the table has an even number of entry slots per
// point, so every point starts 16-byte aligned (sgp_permute_replay_padded).  S slots, the first `nreal` (S or S - 1) of
// them real: half the shared-memory wavefronts of the 8-byte reads (18 -> 10 per pass at d = 8).
template <int VEC, bool FAST, int S>
__device__ __forceinline__ void slice_batch_pairs(const int4 *ep, int nreal, const float *__restrict__ values, int L, int c0,
                                                  Vec<VEC> &acc, float divisor, float rdivisor)
{
    int idx[S];
    float w[S];
    Vec<VEC> v[S];
#pragma unroll
    for (int j = 0; j < S / 2; ++j) {
        const int4 q = ep[j];
        idx[2 * j] = q.x; w[2 * j] = __int_as_float(q.y);
        idx[2 * j + 1] = q.z; w[2 * j + 1] = __int_as_float(q.w);
    }
    int lowest = idx[0];
#pragma unroll
    for (int b = 1; b < S; ++b) lowest = min(lowest, idx[b]);
    if (lowest >= 0) {
#pragma unroll
        for (int b = 0; b < S; ++b)
            if (b < S - 1 || nreal == S) v[b].load_ordered(values + (int64_t)idx[b] * L + c0);
#pragma unroll
        for (int b = 0; b < S; ++b) {
            if (b < S - 1 || nreal == S) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc.v[k] = FAST ? __fmaf_rn(w[b], v[b].v[k], acc.v[k])
                                    : __fadd_rn(acc.v[k], exact_div(__fmul_rn(w[b], v[b].v[k]), divisor, rdivisor));
            }
        }
    }
}

// Persistent warps; warp w takes the tiles w, w + W, ... of P = ppp * passes points.  A tile's replay entries
// ([P, d+1] {index, weight}: contiguous) arrive in the warp's ring by one bulk copy; in a pass the lanes are
// (point, channel chunk): ppp = 32 / chunks points.
// EPI (the CG form, sgp_slice_ring_cg): the sweep that follows the product in a CG iteration -- AP = s * KP + noise * P,
// pAp[l] = sum_n P * AP (sgp_cg_apply) -- runs in the epilogue: the point's row of P is read (coalesced), AP is stored
// instead of KP, and the per-thread dot products are combined per CTA in a fixed order into epi.partial[blockIdx.x, :]
// (a one-block second stage sums those).  Saves a launch and two passes over [N, L] per iteration.
struct SliceEpilogue {
    const float *p;          // P [N, ldp]: the operand of the product
    int64_t ldp;
    const float *s, *noise;  // device scalars
    float *partial;          // [gridDim.x, L]
};

template <int VEC, bool FAST, bool RAGGED, bool EPI = false>
__global__ void __launch_bounds__(RING_THREADS, EPI ? 3 : 0)   // (0 = unspecified: the plain form takes 80 registers by itself)
sgp_slice_ring_kernel(const int2 *__restrict__ replay, const float *__restrict__ values, int64_t N, int dp1, int estride,
                      int L, int chunks, int ppp, int passes, int stages, uint32_t tile_stride, float divisor,
                      float rdivisor, float *__restrict__ out, int64_t ldo, int L_out, SliceEpilogue epi)
{
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = ppp * passes;
    unsigned char *ring = ring_smem + (size_t)warp * stages * tile_stride;
    uint64_t *bars = (uint64_t *)(ring_smem + (size_t)RING_WARPS * stages * tile_stride) + warp * RING_MAX_STAGES;
    const int64_t n_tiles = (N + P - 1) / P;
    const int64_t gw = (int64_t)blockIdx.x * RING_WARPS + warp, W = (int64_t)gridDim.x * RING_WARPS;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    const uint64_t pol = l2_policy_evict_first();

    auto issue = [&](int64_t t, int s) {   // lane 0: stream tile t into stage s
        const int64_t p0 = t * P;
        const int np = (int)min((int64_t)P, N - p0);
        const uint32_t bytes = (uint32_t)np * (uint32_t)estride * 8u;
        const uint32_t b16 = bytes & ~15u;
        unsigned char *dst = ring + (size_t)s * tile_stride;
        const unsigned char *src = (const unsigned char *)(replay + p0 * estride);
        if (bytes != b16) *(int2 *)(dst + b16) = __ldg((const int2 *)(src + b16));   // odd entry count: the last by hand
        mbar_arrive_expect_tx(bars + s, b16);
        bulk_g2s(dst, src, b16, bars + s, pol);
    };
    if (lane == 0) {
        for (int k = 0; k < stages; ++k) {
            const int64_t t = gw + (int64_t)k * W;
            if (t < n_tiles) issue(t, k);
        }
    }
    const int sub = lane / chunks;
    const int c0 = (lane - sub * chunks) * VEC;
    const bool lane_on = sub < ppp && (RAGGED || c0 < L_out);   // (out narrower than the lattice rows by whole vectors)
    Vec<VEC> dot;
    vec_zero(dot);
    float epi_s = 0.0f, epi_noise = 0.0f;
    if (EPI) { epi_s = __ldg(epi.s); epi_noise = __ldg(epi.noise); }
    pdl_wait();   // everything above read build-time tables only; the lattice values are the predecessor's output

    int s = 0;
    uint32_t phase = 0;
    for (int64_t t = gw; t < n_tiles; t += W) {
        mbar_wait(bars + s, phase);
        const int2 *E = (const int2 *)(ring + (size_t)s * tile_stride);
        const int64_t p0 = t * P;
        const int np = (int)min((int64_t)P, N - p0);
        for (int pass = 0; pass < passes; ++pass) {
            const int lp = pass * ppp + sub;
            if (lane_on && lp < np) {
                const int2 *ep = E + lp * estride;
                Vec<VEC> acc;
                vec_zero(acc);
                int r0 = 0;
                if ((estride & 1) == 0) {   // 16-byte aligned points: entries in pairs
                    for (; r0 + 10 <= estride; r0 += 10)
                        slice_batch_pairs<VEC, FAST, 10>((const int4 *)(ep + r0), min(10, dp1 - r0), values, L, c0, acc, divisor, rdivisor);
                    for (; r0 + 4 <= estride; r0 += 4)
                        slice_batch_pairs<VEC, FAST, 4>((const int4 *)(ep + r0), min(4, dp1 - r0), values, L, c0, acc, divisor, rdivisor);
                    for (; r0 + 2 <= estride; r0 += 2)
                        slice_batch_pairs<VEC, FAST, 2>((const int4 *)(ep + r0), min(2, dp1 - r0), values, L, c0, acc, divisor, rdivisor);
                } else {
                    for (; r0 + 9 <= dp1; r0 += 9) slice_batch<VEC, FAST, 9>(ep + r0, values, L, c0, acc, divisor, rdivisor);
                    for (; r0 + 3 <= dp1; r0 += 3) slice_batch<VEC, FAST, 3>(ep + r0, values, L, c0, acc, divisor, rdivisor);
                    for (; r0 < dp1; ++r0) slice_batch<VEC, FAST, 1>(ep + r0, values, L, c0, acc, divisor, rdivisor);
                }
                if (FAST) {   // one division of the sum instead of one per term (differs from the reference by rounding only)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc.v[k] = exact_div(acc.v[k], divisor, rdivisor);
                }
                float *orow = out + (p0 + lp) * ldo + c0;
                if (EPI) {
                    Vec<VEC> pv;   // (requesting it ahead of the row gathers measured slower: 16 more bytes of spills)
                    pv.load_plain(epi.p + (p0 + lp) * epi.ldp + c0);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        acc.v[k] = fmaf(epi_s, acc.v[k], epi_noise * pv.v[k]);
                        dot.v[k] = fmaf(pv.v[k], acc.v[k], dot.v[k]);
                    }
                    acc.store(orow);   // read again by the next sweep: no streaming hint
                } else if (RAGGED) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (c0 + k < L_out) __stcs(orow + k, acc.v[k]);
                } else {
                    acc.store_streaming(orow);
                }
            }
        }
        __syncwarp();   // every lane is done with stage s
        const int64_t tn = t + (int64_t)stages * W;
        if (lane == 0 && tn < n_tiles) issue(tn, s);
        if (++s == stages) { s = 0; phase ^= 1u; }
    }
    if (EPI) {
        // per CTA and column: the threads' dot products in thread order (deterministic: the tile assignment is static)
        __shared__ float s_dot[VEC][RING_THREADS];
#pragma unroll
        for (int k = 0; k < VEC; ++k) s_dot[k][threadIdx.x] = lane_on ? dot.v[k] : 0.0f;
        __syncthreads();
        if ((int)threadIdx.x < L_out) {
            const int chunk = threadIdx.x / VEC, kk = threadIdx.x % VEC;
            float t = 0.0f;
            for (int w = 0; w < RING_WARPS; ++w)
                for (int sb = 0; sb < ppp; ++sb) t += s_dot[kk][w * 32 + sb * chunks + chunk];
            epi.partial[(int64_t)blockIdx.x * L_out + threadIdx.x] = t;
        }
    }
}

// =====================================================================================================
// splat
// =====================================================================================================
// Tile geometry shared by the zero kernel and the splat: a pass covers spp = 32 / chunks segments of 8 entries.
struct SplatTile {
    int spp, passes, T;   // T = 8 * spp * passes entries per tile, a multiple of 64 (whole interleave groups)
};
static SplatTile splat_tile(int chunks, int target)
{
    SplatTile g;
    g.spp = 32 / chunks;
    if (g.spp < 1) g.spp = 1;
    int passes = target / (8 * g.spp);
    if (passes < 1) passes = 1;
    while ((g.spp * passes) & 7) ++passes;   // T % 64 == 0: whole interleave groups
    g.passes = passes;
    g.T = 8 * g.spp * passes;
    return g;
}

// A lattice row whose entries continue across a tile boundary is accumulated with reductions: zero it first.
// One thread group per tile t >= 1 whose first entry does not start a row.
__global__ void __launch_bounds__(256)
sgp_ring_zero_heads_kernel(const int2 *__restrict__ ent, const int32_t *__restrict__ seg_row, int64_t n_tiles, int T,
                           int L, float *__restrict__ values)
{
    pdl_launch_dependents();
    pdl_wait();   // the previous product may still be reading this buffer
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = tid / L + 1;
    if (t >= n_tiles) return;
    const int c = (int)(tid - (t - 1) * L);
    const int64_t e0 = t * (int64_t)T;
    if (__ldg(&ent[e0].x) < 0) return;
    values[(int64_t)__ldg(seg_row + e0 / 4) * L + c] = 0.0f;
}

// A tile is a whole number of interleave groups (64 entries, sgp_entry_index): one bulk copy brings it in as it lies in
// global memory, and the 16-byte pieces that the segments of a pass read together are contiguous in shared memory too
// (no bank conflict).  ent_offset(e) = byte offset of row-sorted entry e inside the tile.
__device__ __forceinline__ uint32_t ent_offset(int entry) { return (uint32_t)sgp_entry_index(entry) * 8u; }

// thread = (segment of 8 consecutive row-sorted entries, channel chunk) within a pass; see the file header.
// RAGGED: src has L_src < L columns / arbitrary alignment and is read channel by channel (missing channels = 0).
// SCAN: combine the runs of a tile across its threads and store (values not memset, boundary rows zeroed by
//       sgp_ring_zero_heads_kernel); otherwise one reduction per run and thread into memset values.
template <int VEC, bool RAGGED, bool SCAN>
__global__ void __launch_bounds__(RING_THREADS, SCAN ? 3 : 4)
sgp_splat_ring_kernel(const int2 *__restrict__ ent, const int32_t *__restrict__ seg_row, int64_t n_entries,
                      const float *__restrict__ src, int64_t lds, int L, int L_src, int chunks, int live, int spp, int passes,
                      int stages, uint32_t tile_stride, float *__restrict__ values)
{
    // chunks = lane slots per segment (the lane layout and the shuffle strides), live <= chunks = the channel chunks that
    // exist: with 3 (5, 6, 7) chunks per row the segments still sit at power-of-two lane strides and the spare lanes idle,
    // which keeps the warp-uniform aggregation below (idle lanes cost no memory wavefronts, and issue slots are not the limit)
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = 8 * spp * passes;
    const uint32_t ent_region = (uint32_t)(T + 2) * 8u;
    unsigned char *ring = ring_smem + (size_t)warp * stages * tile_stride;
    uint64_t *bars = (uint64_t *)(ring_smem + (size_t)RING_WARPS * stages * tile_stride) + warp * RING_MAX_STAGES;
    const int64_t n_tiles = (n_entries + T - 1) / T;
    const int64_t gw = (int64_t)blockIdx.x * RING_WARPS + warp, W = (int64_t)gridDim.x * RING_WARPS;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    const uint64_t pol = l2_policy_evict_first();

    // lane 0: the tile's entries (+ the first entry pair of the next tile: the look-ahead flag) and its row ids
    auto issue = [&](int64_t t, int s) {
        if (lane != 0) return;
        const int64_t e0 = t * (int64_t)T;
        const uint32_t ne = (uint32_t)min((int64_t)T, n_entries - e0);
        const uint32_t look = (SCAN && e0 + ne < n_entries) ? 16u : 0u;
        unsigned char *dst = ring + (size_t)s * tile_stride;
        mbar_arrive_expect_tx(bars + s, ne * 8u + look + ne);
        bulk_g2s(dst, ent + e0, ne * 8u + look, bars + s, pol);
        bulk_g2s(dst + ent_region, seg_row + e0 / 4, ne, bars + s, pol);
    };
    for (int k = 0; k < stages; ++k) {
        const int64_t t = gw + (int64_t)k * W;
        if (t < n_tiles) issue(t, k);
    }
    const int sub = lane / chunks;
    const int cl = lane - sub * chunks;
    const int c0 = cl * VEC;
    const bool lane_on = sub < spp && cl < live;
    const int last_src = (spp - 1) * chunks + cl;   // the lane holding this chunk of the pass's last segment
    pdl_wait();   // values is zeroed (SCAN: its boundary rows are) by the stream's previous work

    int s = 0;
    uint32_t phase = 0;
    for (int64_t t = gw; t < n_tiles; t += W) {
        mbar_wait(bars + s, phase);
        const unsigned char *E = ring + (size_t)s * tile_stride;
        const int32_t *R = (const int32_t *)(E + ent_region);
        const int64_t e0 = t * (int64_t)T;
        const int ne = (int)min((int64_t)T, n_entries - e0);
        const int nseg = ne >> 3;
        const bool has_look = e0 + ne < n_entries;
        const bool head_open = SCAN && t > 0 && *(const int *)E >= 0;   // the tile's first entry continues a row of the previous tile
        Vec<VEC> C;                                               // open run carried from pass to pass
        vec_zero(C);
        bool Cr = false;                                          // a row start was seen since the tile began
        for (int pass = 0; pass < passes; ++pass) {
            const int seg = pass * spp + sub;
            const bool act = lane_on && seg < nseg;
            Vec<VEC> tail, p0v;
            vec_zero(tail);
            vec_zero(p0v);
            bool reset = false, f0 = false, next_flag = true;
            int k = 0, row0 = 0;
            if (act) {
                const int4 *ep = (const int4 *)(E + ent_offset(seg * 8));
                int pt[8];
                float w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int4 q = ep[8 * i];   // piece i of this segment (the 8 segments of a group are interleaved)
                    pt[2 * i] = q.x; w[2 * i] = __int_as_float(q.y);
                    pt[2 * i + 1] = q.z; w[2 * i + 1] = __int_as_float(q.w);
                }
                row0 = R[seg * 2];
                if (SCAN && (seg + 1 < nseg || has_look)) next_flag = *(const int *)(E + ent_offset((seg + 1) * 8)) < 0;
                // no weight has all bits set and rows are never negative: the branch keeps every shared-memory read
                // ahead of every row load (warps issue in order)
                int all = __float_as_int(w[0]);
#pragma unroll
                for (int i = 1; i < 8; ++i) all &= __float_as_int(w[i]);
                if (all != -1 && row0 >= 0) {
                    Vec<VEC> v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float *p = src + (int64_t)(pt[i] & 0x7fffffff) * lds + c0;
                        if (RAGGED) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) v[i].v[q] = (c0 + q < L_src) ? ldg_ordered_f1(p + q) : 0.0f;
                        } else {
                            v[i].load_ordered(p);
                        }
                    }
                    f0 = pt[0] < 0 || (t == 0 && seg == 0);   // (the very first entry starts a row but carries no flag)
                    Vec<VEC> acc;
                    vec_zero(acc);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) acc.v[q] = __fmaf_rn(w[i], v[i].v[q], acc.v[q]);
                        if (i < 7 && pt[i + 1] < 0) {   // the next entry starts the next lattice row: this piece is closed
                            if (!SCAN) acc.red(values + (int64_t)(row0 + k) * L + c0);
                            else if (k == 0) p0v = acc;   // the first piece may continue a run of earlier threads: emitted below
                            else acc.store(values + (int64_t)(row0 + k) * L + c0);
                            ++k;
                            vec_zero(acc);
                        }
                    }
                    tail = acc;
                    reset = (k > 0) || f0;
                }
            }
            if (!SCAN) {
                // Reductions: one per run and thread -- except in the warp-uniform case, where every segment of the pass
                // lies inside ONE lattice row (the long rows at the centre of the data, where most reductions go): the
                // partial sums are added across the segments with a butterfly and one lane group reduces (8x fewer L2
                // atomics for those passes; every other pass pays one vote).
                const bool pow2 = (chunks & (chunks - 1)) == 0;
                const int row_first = __shfl_sync(0xffffffffu, row0, 0);
                const bool uniform = pow2 && __all_sync(0xffffffffu, cl >= live || (act && k == 0 && row0 == row_first));
                if (uniform) {
                    for (int step = chunks; step < 32; step <<= 1) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) tail.v[q] += __shfl_xor_sync(0xffffffffu, tail.v[q], step);
                    }
                    if (lane < live) tail.red(values + (int64_t)row0 * L + c0);
                } else if (act) {
                    tail.red(values + (int64_t)(row0 + k) * L + c0);
                }
                continue;
            }
            // inclusive segmented scan of (tail, reset) over the segments of the pass, per channel chunk
            Vec<VEC> sv = tail;
            bool sr = reset;
            for (int off = chunks; off < 32; off <<= 1) {
                const int pr = __shfl_up_sync(0xffffffffu, (int)sr, off);
                Vec<VEC> pv;
#pragma unroll
                for (int q = 0; q < VEC; ++q) pv.v[q] = __shfl_up_sync(0xffffffffu, sv.v[q], off);
                if (lane >= off) {
                    if (!sr) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) sv.v[q] += pv.v[q];
                    }
                    sr = sr || (pr != 0);
                }
            }
            // what the earlier segments of this pass (and the earlier passes) carry into this segment
            Vec<VEC> ev;
            int er = __shfl_up_sync(0xffffffffu, (int)sr, chunks);
#pragma unroll
            for (int q = 0; q < VEC; ++q) ev.v[q] = __shfl_up_sync(0xffffffffu, sv.v[q], chunks);
            if (lane < chunks) {
                vec_zero(ev);
                er = 0;
            }
            Vec<VEC> cin;
#pragma unroll
            for (int q = 0; q < VEC; ++q) cin.v[q] = er ? ev.v[q] : C.v[q] + ev.v[q];
            const bool cin_from_head = head_open && !Cr && !er;   // the carried run began before this tile
            if (act) {
                if (k > 0) {   // first piece, closed inside the segment
                    if (!f0) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) p0v.v[q] += cin.v[q];
                    }
                    float *dst = values + (int64_t)row0 * L + c0;
                    if (!f0 && cin_from_head) p0v.red(dst); else p0v.store(dst);
                }
                const bool joins = (k == 0) && !f0;   // the last piece is the first piece and continues the carried run
                const bool tile_end = (seg == nseg - 1);
                if (next_flag || tile_end) {
                    Vec<VEC> o = tail;
                    if (joins) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) o.v[q] += cin.v[q];
                    }
                    float *dst = values + (int64_t)(row0 + k) * L + c0;
                    if (!next_flag || (joins && cin_from_head)) o.red(dst); else o.store(dst);
                }
            }
            // carry to the next pass
            const int lr = __shfl_sync(0xffffffffu, (int)sr, last_src);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const float lv = __shfl_sync(0xffffffffu, sv.v[q], last_src);
                C.v[q] = lr ? lv : C.v[q] + lv;
            }
            Cr = Cr || (lr != 0);
        }
        __syncwarp();   // every lane is done with stage s
        const int64_t tn = t + (int64_t)stages * W;
        if (tn < n_tiles) issue(tn, s);
        if (++s == stages) { s = 0; phase ^= 1u; }
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
static int ring_env(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// read on every call (a getenv, ~100 ns): a process can switch between the two forms, e.g. to compare them
extern "C" int sgp_ring_enabled(void) { return ring_env("SGP_RING", 1) != 0; }
// The ring splat: first measured slower than the one-shot kernel (93 vs 86 us at the metric shape) with entries stored
// linearly (18 bulk copies per tile and padded shared-memory blocks); with the interleaved entry layout (one bulk copy per
// tile, conflict-free reads) and the warp-uniform aggregation of long rows it measures 74.7 us and the MVM 178.5 us
// (188.8 with the one-shot splat).  sgp_splat_rows selects it for dense lattices only (see there).
extern "C" int sgp_ring_splat_enabled(void) { return sgp_ring_enabled() && ring_env("SGP_RING_SPLAT", 1) != 0; }
extern "C" int sgp_ring_slice_enabled(void) { return sgp_ring_enabled() && ring_env("SGP_RING_SLICE", 1) != 0; }

struct RingLaunch {
    int stages;
    uint32_t tile_stride;
    size_t smem;
    unsigned grid;
};

// persistent grid: as many CTAs as fit on the device at this shared-memory size, at most one warp per tile
template <typename K>
static int ring_config(K kernel, uint32_t tile_bytes, int64_t n_tiles, int stages_dflt, const char *stages_env,
                       RingLaunch *rl)
{
    int stages = ring_env(stages_env, stages_dflt);
    if (stages < 1) stages = 1;
    if (stages > RING_MAX_STAGES) stages = RING_MAX_STAGES;
    rl->stages = stages;
    rl->tile_stride = (tile_bytes + 127u) & ~127u;
    rl->smem = (size_t)RING_WARPS * stages * rl->tile_stride + (size_t)RING_WARPS * RING_MAX_STAGES * sizeof(uint64_t);
    if (rl->smem > 227 * 1024) return fail(SGP_EUNSUPPORTED, "ring kernel needs %zu bytes of shared memory", rl->smem);
    // the opt-in is per device and cheap: set it on every launch rather than caching it per process
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl->smem));
    int dev = 0, sms = 0, occ = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, RING_THREADS, rl->smem));
    if (occ < 1) return fail(SGP_EUNSUPPORTED, "ring kernel does not fit an SM (%zu bytes of shared memory)", rl->smem);
    const int occ_cap = ring_env("SGP_RING_OCC", 0);
    if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
    int64_t grid = (int64_t)sms * occ;
    const int64_t need = (n_tiles + RING_WARPS - 1) / RING_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    rl->grid = (unsigned)grid;
    return SGP_OK;
}

// widest vector the lattice rows allow; 0 if the channel chunks do not fit a warp
static int ring_vec(int L, const void *rows)
{
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    int vec = 1;
    if (L % 4 == 0 && al(rows, 16)) vec = 4;
    else if (L % 2 == 0 && al(rows, 8)) vec = 2;
    return (L / vec <= 32) ? vec : 0;
}

// points per pass and passes per tile of the slice; 0 if the shape does not fit a warp's ring
static int slice_geometry(int dp1, int chunks, int *ppp_out, int *passes_out)
{
    if (chunks < 1 || chunks > 32) return 0;
    const int ppp = 32 / chunks;
    // tiles are multiples of 16 bytes (ppp * passes * dp1 even) and at most 2.25 KB: the kernel's speed is proportional
    // to the resident warps (83 / 64 / 52 us at 24 / 32 / 40 warps per SM at the metric shape), and 2 stages x 8 warps x
    // 2.25 KB still leaves shared memory for 6 CTAs; at L = 16 two to four passes of 8 points measure alike kernel-only
    // (52.7 / 52.0 / 52.1 us), three or four are 1.5 us better inside the MVM's graph (172.2 -> 170.7 us)
    int passes = ring_env("SGP_SLICE_PASSES", 3);
    if (passes < 1) passes = 1;
    if ((ppp * passes * dp1) & 1) ++passes;
    while (passes > 1 && (size_t)ppp * passes * dp1 * 8 > 2304) {
        --passes;
        if ((ppp * passes * dp1) & 1) --passes;
    }
    if (passes < 1) passes = ((ppp * dp1) & 1) ? 2 : 1;
    if ((size_t)ppp * passes * dp1 * 8 > 3072) return 0;   // larger stages cost more occupancy than the prefetch gains: one-shot kernel
    *ppp_out = ppp;
    *passes_out = passes;
    return 1;
}

extern "C" int sgp_slice_ring_supported(const sgp_lattice_view *lat, const float *values, int L)
{
    if (!lat || lat->perm || lat->replay_transposed) return 0;
    const int vec = ring_vec(L, values);
    int ppp, passes;
    const int estride = lat->replay_stride > 0 ? lat->replay_stride : lat->d + 1;
    if (estride > lat->d + 2) return 0;   // at most one filler slot per point
    if (vec == 0 || !slice_geometry(estride, L / vec, &ppp, &passes)) return 0;
    // narrow rows: with one or two channel chunks a warp instruction gathers 32 / 16 distinct rows and the one-shot
    // kernel measures faster (L = 1 / 2 / 4 / 8: 28 / 33 / 38 / 45 us against 64 / 67 / 45 / 49 us; L = 12 / 16 / 32: 59 / 62 / 110 against 57 / 51 / 96); SGP_RING_FORCE=1 overrides (experiments)
    return L / vec >= ring_env("SGP_RING_MIN_CHUNKS", 3) || ring_env("SGP_RING_FORCE", 0) != 0;
}

extern "C" int sgp_slice_ring(const sgp_lattice_view *lat, const float *values, int L, float *out, int64_t ldo,
                              int L_out, sgp_stream_t stream)
{
    SGP_RANGE("sgp_slice_ring");
    if (!lat || lat->N < 0 || lat->d < 1 || lat->d > SGP_MAX_DIM || L < 1) return fail(SGP_EINVAL, "sgp_slice_ring: bad view");
    if (lat->N == 0) return SGP_OK;
    if (!values || !out || !lat->replay || L_out < 1 || L_out > L || ldo < L_out)
        return fail(SGP_EINVAL, "sgp_slice_ring: null pointer, L_out outside [1, L] or ldo < L_out");
    if (!sgp_slice_ring_supported(lat, values, L)) return fail(SGP_EUNSUPPORTED, "sgp_slice_ring: shape not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = ring_vec(L, values);
    const bool ragged = vec > 1 && !(L_out % vec == 0 && ldo % vec == 0 && ((uintptr_t)out % (4 * vec)) == 0);
    if (vec == 1 && L_out != L) return fail(SGP_EUNSUPPORTED, "sgp_slice_ring: L_out < L needs vectorisable lattice rows");
    const int chunks = L / vec;
    const int dp1 = lat->d + 1;
    const int estride = lat->replay_stride > 0 ? lat->replay_stride : dp1;
    int ppp = 0, passes = 0;
    slice_geometry(estride, chunks, &ppp, &passes);
    const int P = ppp * passes;
    const int64_t n_tiles = (lat->N + P - 1) / P;
    const float divisor = sgp_slice_divisor(lat->d);
    volatile float rdivisor = 1.0f / divisor;
    RingLaunch rl;
    int rc;
    cudaError_t le = cudaSuccess;
#define SGP_SLICE_RING(VV, FF, RG)                                                                                     \
    do {                                                                                                               \
        rc = ring_config(sgp_slice_ring_kernel<VV, FF, RG>, (uint32_t)P * estride * 8u, n_tiles, 2, "SGP_SLICE_STAGES", &rl); \
        if (rc) return rc;                                                                                             \
        le = sgp_launch_pdl(sgp_slice_ring_kernel<VV, FF, RG>, dim3(rl.grid), dim3(RING_THREADS), rl.smem, st,         \
                            (const int2 *)lat->replay, values, lat->N, dp1, estride, L, chunks, ppp, passes, rl.stages, \
                            rl.tile_stride, divisor, (float)rdivisor, out, ldo, L_out, SliceEpilogue{});               \
    } while (0)
#define SGP_SLICE_RING_V(VV)                                                                                           \
    do {                                                                                                               \
        if (lat->fast) { if (ragged) SGP_SLICE_RING(VV, true, true); else SGP_SLICE_RING(VV, true, false); }           \
        else { if (ragged) SGP_SLICE_RING(VV, false, true); else SGP_SLICE_RING(VV, false, false); }                   \
    } while (0)
    if (vec == 4) SGP_SLICE_RING_V(4);
    else if (vec == 2) SGP_SLICE_RING_V(2);
    else SGP_SLICE_RING_V(1);
#undef SGP_SLICE_RING_V
#undef SGP_SLICE_RING
    if (le != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_slice_ring_kernel failed: %s", cudaGetErrorString(le));
    return launch_ok("sgp_slice_ring_kernel");
}

// The CG form: out = s * slice(values) + noise * P, pAp[l] = sum_n P[n, l] * out[n, l] (sgp_cg_apply fused into the
// slice).  16-byte vectors only: L % 4 == 0, out / P aligned with ldo, ldp % 4 == 0.  scratch: sgp_cg_scratch_floats(L).
extern "C" int sgp_slice_ring_cg_supported(const sgp_lattice_view *lat, const float *values, int L, const float *out,
                                           int64_t ldo, int L_out, const float *P, int64_t ldp)
{
    auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    return sgp_ring_slice_enabled() && sgp_slice_ring_supported(lat, values, L) && ring_vec(L, values) == 4 && al16(out) &&
           al16(P) && ldo % 4 == 0 && ldp % 4 == 0 && L_out >= 4 && L_out <= L && L_out % 4 == 0 && ldo >= L_out &&
           ldp >= L_out && L_out <= RING_THREADS;
}

extern "C" int sgp_slice_ring_cg(const sgp_lattice_view *lat, const float *values, int L, float *out, int64_t ldo,
                                 int L_out, const float *P, int64_t ldp, const float *s, const float *noise, float *pAp,
                                 float *scratch, sgp_stream_t stream)
{
    SGP_RANGE("sgp_slice_ring_cg");
    if (!lat || lat->N < 0 || lat->d < 1 || lat->d > SGP_MAX_DIM || L < 1) return fail(SGP_EINVAL, "sgp_slice_ring_cg: bad view");
    if (!values || !out || !P || !s || !noise || !pAp || !scratch || !lat->replay)
        return fail(SGP_EINVAL, "sgp_slice_ring_cg: null pointer");
    if (lat->N == 0) return SGP_OK;
    if (!sgp_slice_ring_cg_supported(lat, values, L, out, ldo, L_out, P, ldp))
        return fail(SGP_EUNSUPPORTED, "sgp_slice_ring_cg: shape not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = L / 4;
    const int dp1 = lat->d + 1;
    const int estride = lat->replay_stride > 0 ? lat->replay_stride : dp1;
    int ppp = 0, passes = 0;
    slice_geometry(estride, chunks, &ppp, &passes);
    const int P_tile = ppp * passes;
    const int64_t n_tiles = (lat->N + P_tile - 1) / P_tile;
    const float divisor = sgp_slice_divisor(lat->d);
    volatile float rdivisor = 1.0f / divisor;
    RingLaunch rl;
    int rc;
    cudaError_t le = cudaSuccess;
    SliceEpilogue epi{P, ldp, s, noise, scratch};
#define SGP_SLICE_RING_CG(FF)                                                                                          \
    do {                                                                                                               \
        rc = ring_config(sgp_slice_ring_kernel<4, FF, false, true>, (uint32_t)P_tile * estride * 8u, n_tiles, 2,       \
                         "SGP_SLICE_STAGES", &rl);                                                                     \
        if (rc) return rc;                                                                                             \
        if ((size_t)rl.grid * (size_t)L_out > sgp_cg_scratch_floats(L_out))                                            \
            return fail(SGP_EUNSUPPORTED, "sgp_slice_ring_cg: %u CTAs exceed the scratch", rl.grid);                   \
        le = sgp_launch_pdl(sgp_slice_ring_kernel<4, FF, false, true>, dim3(rl.grid), dim3(RING_THREADS), rl.smem, st, \
                            (const int2 *)lat->replay, values, lat->N, dp1, estride, L, chunks, ppp, passes, rl.stages, \
                            rl.tile_stride, divisor, (float)rdivisor, out, ldo, L_out, epi);                           \
    } while (0)
    if (lat->fast) SGP_SLICE_RING_CG(true); else SGP_SLICE_RING_CG(false);
#undef SGP_SLICE_RING_CG
    if (le != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_slice_ring_kernel (CG form) failed: %s", cudaGetErrorString(le));
    rc = launch_ok("sgp_slice_ring_kernel");
    if (rc) return rc;
    return sgp_cg_reduce_partials(scratch, (int)rl.grid, L_out, pAp, stream);
}

extern "C" int sgp_splat_ring_supported(const float *values, int L) { return ring_vec(L, values) != 0; }

static int splat_rows_ring_impl(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                const float *src, int64_t lds, int L_src, float *values, int L, bool prezeroed,
                                sgp_stream_t stream);

extern "C" int sgp_splat_rows_ring(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                   const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream)
{
    return splat_rows_ring_impl(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, false, stream);
}

// values already holds zeros (the caller zeroed it off the critical path); reductions form only
int sgp_splat_rows_ring_prezeroed(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                  const float *src, int64_t lds, int L_src, float *values, int L, sgp_stream_t stream)
{
    return splat_rows_ring_impl(ent, seg_row, n_entries, N, M, src, lds, L_src, values, L, true, stream);
}

static int splat_rows_ring_impl(const int32_t *ent, const int32_t *seg_row, int64_t n_entries, int64_t N, int64_t M,
                                const float *src, int64_t lds, int L_src, float *values, int L, bool prezeroed,
                                sgp_stream_t stream)
{
    SGP_RANGE("sgp_splat_rows_ring");
    if (N == 0 || M == 0) return SGP_OK;
    if (!ent || !seg_row || !src || !values || N < 0 || M < 0 || n_entries < 64 || n_entries % 64 != 0 || L_src < 1 ||
        lds < L_src || L < L_src)
        return fail(SGP_EINVAL, "sgp_splat_rows_ring: bad argument");
    const int vec = ring_vec(L, values);
    if (!vec) return fail(SGP_EUNSUPPORTED, "sgp_splat_rows_ring: %d channels do not fit a warp", L);
    cudaStream_t st = (cudaStream_t)stream;
    auto al = [](const void *p, int bytes) { return ((uintptr_t)p % bytes) == 0; };
    // src narrower than the lattice rows by whole vectors (a 12-column block on 16-channel lattice rows, which keeps every
    // row gather inside one 128-byte line): the spare lane slots idle, no channel-by-channel path needed
    const bool whole_chunks = L_src % vec == 0 && lds % vec == 0 && al(src, 4 * vec);
    const bool ragged = !whole_chunks;
    const int live = whole_chunks ? L_src / vec : L / vec;
    RingLaunch rl;
    int rc;
    cudaError_t le = cudaSuccess;
    // SGP_SPLAT_SCAN=1: runs combined across the threads of a tile and stored, no memset (the shuffles of the scan cost
    // as many L1 data-pipe wavefronts as a third of the row gathers: measured slower at the metric shape, 93 vs 8x us)
    const bool scan = !prezeroed && ring_env("SGP_SPLAT_SCAN", 0) != 0 && live == L / vec;   // (the store form writes every channel)
    // lane slots per segment: rounded up to a power of two in the reductions form (see the kernel), SGP_SPLAT_SLOTS=0: not
    int chunks = L / vec;
    if (!scan && chunks <= 16 && ring_env("SGP_SPLAT_SLOTS", 1) != 0)
        while (chunks & (chunks - 1)) ++chunks;
    // Entries per tile.  A persistent warp takes the tiles w, w + W, ...: with 256-entry tiles the metric shape has 7.4
    // tiles per warp and the last round runs with 42 % of the warps (67.9 us); 192-entry tiles make it 9.9 (61.9 us),
    // 128: 14.9 (63.3 us), 64: 70 us (profiles/exp_splat_tile.py).  Pick the size whose last round is fullest, smaller
    // tiles paying a little for their extra bulk copies and waits.  SGP_SPLAT_TILE fixes it.
    int target = ring_env("SGP_SPLAT_TILE", 0);
    if (target < 64) {
        int dev = 0, sms = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const double W = (double)sms * 4 * RING_WARPS;   // 4 CTAs per SM (registers)
        const int cand[3] = {192, 256, 128};
        const double penalty[3] = {0.005, 0.0, 0.02};
        double best = -1.0;
        target = 256;
        for (int c = 0; c < 3; ++c) {
            const SplatTile t = splat_tile(chunks, cand[c]);
            const double rounds = (double)((n_entries + t.T - 1) / t.T) / W;
            const double fill = rounds <= 1.0 ? 1.0 : rounds / (double)(int64_t)(rounds + 0.999999);
            if (fill - penalty[c] > best) { best = fill - penalty[c]; target = cand[c]; }
        }
    }
    const SplatTile g = splat_tile(chunks, target);
    const int64_t n_tiles = (n_entries + g.T - 1) / g.T;
    if (!scan) {
        if (!prezeroed) CUDA_TRY(cudaMemsetAsync(values, 0, sizeof(float) * (size_t)M * (size_t)L, st));
    } else if (n_tiles > 1) {
        const int64_t work = (n_tiles - 1) * L;
        le = sgp_launch_pdl(sgp_ring_zero_heads_kernel, dim3(sgp_grid_for(work, 256)), dim3(256), 0, st,
                            (const int2 *)ent, seg_row, n_tiles, g.T, L, values);
        if (le != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_ring_zero_heads_kernel failed: %s", cudaGetErrorString(le));
    }
    const uint32_t tile_bytes = (uint32_t)(g.T + 2) * 8u + (uint32_t)g.T;
#define SGP_SPLAT_RING(VV, RG, SC)                                                                                     \
    do {                                                                                                               \
        rc = ring_config(sgp_splat_ring_kernel<VV, RG, SC>, tile_bytes, n_tiles, 2, "SGP_SPLAT_STAGES", &rl);          \
        if (rc) return rc;                                                                                             \
        le = sgp_launch_pdl(sgp_splat_ring_kernel<VV, RG, SC>, dim3(rl.grid), dim3(RING_THREADS), rl.smem, st,         \
                            (const int2 *)ent, seg_row, n_entries, src, lds, L, L_src, chunks, live, g.spp, g.passes,  \
                            rl.stages, rl.tile_stride, values);                                                        \
    } while (0)
#define SGP_SPLAT_RING_S(VV, RG) do { if (scan) SGP_SPLAT_RING(VV, RG, true); else SGP_SPLAT_RING(VV, RG, false); } while (0)
    if (vec == 4) { if (ragged) SGP_SPLAT_RING_S(4, true); else SGP_SPLAT_RING_S(4, false); }
    else if (vec == 2) { if (ragged) SGP_SPLAT_RING_S(2, true); else SGP_SPLAT_RING_S(2, false); }
    else { if (ragged) SGP_SPLAT_RING_S(1, true); else SGP_SPLAT_RING_S(1, false); }
#undef SGP_SPLAT_RING_S
#undef SGP_SPLAT_RING
    if (le != cudaSuccess) return fail(SGP_ECUDA, "launch of sgp_splat_ring_kernel failed: %s", cudaGetErrorString(le));
    return launch_ok("sgp_splat_ring_kernel");
}
