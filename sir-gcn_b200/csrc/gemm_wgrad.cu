// Weight (and bias) gradients of the layer's projections on the tcgen05 tensor cores (sm_100a):
//     dW[Nout, Kin] = dY[M, Nout]^T · X[M, Kin]        db[Nout] = Σ_m dY[m, :]
// (autograd backward of nn.Linear behind /root/reference/models/conv.py:60-61,:65).  The
// contraction runs over the NODES (M = 5·10^7 on the headline graph) while the output is a few hundred rows: both
// operands are consumed "MN-major" — a TMA box of 64 nodes x 64 features lands in shared memory as 64 rows of 128
// bytes, which IS the canonical MN-major SWIZZLE_128B operand layout with the node index as the MMA's K dimension —
// so neither table is ever transposed.  dY and X are each read once per step (HBM-bound: 768 B per node for
// d = 128), fp32 accumulation in TMEM.
//   grid = (node splits, column ranges of X, row groups of dW); every CTA accumulates its node range for up to 4
//   MMA-M tiles (512 rows of dW) x `nc` columns in TMEM, writes an fp32 partial, and a second kernel sums the partials
//   of the node splits in split order (no atomics: bitwise repeatable).
//   The bias gradient rides along as one more MMA per step against a constant tile of ones (N = 16): the column sums
//   of dY appear as 16 extra accumulator columns — the separate column-sum pass over dY (11 ms per step on the 2 B-edge
//   graph) disappears.
//   warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue (tcgen05.ld -> partial).
#include "tc_common.cuh"

namespace sirgcn {
namespace {

constexpr int kWgNodes = 64;                    // nodes per ring stage
constexpr int kWgBox = kWgNodes * 128;          // one box: 64 nodes x 128 bytes (64 16-bit features) = 8 KB
constexpr int kWgThreads = 192;
constexpr int kWgMaxMt = 4;                     // MMA-M tiles (128 rows of dW each) per CTA
constexpr int kWgSmemBudget = 200 * 1024;
constexpr int kWgMaxBlocks = 512;               // node blocks accumulated by one CTA (see wg_grid)

struct WgShape {
    int mt;          // M tiles handled per CTA (row group of dW = mt * 128 rows)
    int na;          // dY boxes per stage (= mt * 2, clipped to the table)
    int nc;          // columns of X per CTA (MMA N, multiple of 16, <= 256)
    int nb;          // X boxes per stage = ceil(nc / 64)
    int ncp;         // accumulator columns per M tile (nc, + 16 when the bias gradient rides along)
    int stages;
    int tmem_cols;
};

__device__ __forceinline__ uint32_t umma_idesc_mn(int n, bool bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                                    // D format: F32
    d |= (bf16 ? 1u : 0u) << 7;                      // A format
    d |= (bf16 ? 1u : 0u) << 10;                     // B format
    d |= 1u << 15;                                   // A is MN-major
    d |= 1u << 16;                                   // B is MN-major
    d |= (uint32_t)(n >> 3) << 17;                   // N >> 3
    d |= (uint32_t)(128 >> 4) << 24;                 // M >> 4
    return d;
}

template <bool BF16>
__global__ void __launch_bounds__(kWgThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x,
                  float *__restrict__ partial, int M, int Nout, int Kin, WgShape sh, int with_bias, int ldp,
                  int blocks_per_split) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (sh.mt * 2 + sh.nb) * kWgBox;
    unsigned char *s_ones = smem + sh.stages * stage_bytes;              // 2 KB of 1.0 (16 nodes x 128 B), 1024-aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_ones + 2048);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * 8 + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t done_bar = bar0 + 8u * 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.z * sh.mt * 128;                             // first row of dW (= column of dY) of this CTA
    const int n0 = blockIdx.y * sh.nc;                                   // first column of dW (= column of X)
    const int node_blocks = (M + kWgNodes - 1) / kWgNodes;
    const int blk0 = blockIdx.x * blocks_per_split, blk1 = min(node_blocks, blk0 + blocks_per_split);
    const bool bias_here = with_bias && blockIdx.y == 0;
    // boxes that overlap the tables (a box wholly outside is never loaded: its rows / columns of the result are masked)
    const int na = min(sh.na, (Nout - m0 + 63) / 64), nb = min(sh.nb, (Kin - n0 + 63) / 64);

    if (threadIdx.x == 0) {
        for (int s = 0; s < sh.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    // the ones tile, and zeros in every box the producer will never fill (stale shared memory must not reach the MMA
    // as NaN patterns even though those rows / columns are masked)
    {
        const uint32_t one2 = BF16 ? 0x3f803f80u : 0x3c003c00u;
        for (int i = threadIdx.x; i < 2048 / 4; i += kWgThreads) reinterpret_cast<uint32_t *>(s_ones)[i] = one2;
        for (int s = 0; s < sh.stages; ++s) {
            for (int b = na; b < sh.mt * 2; ++b)
                for (int i = threadIdx.x; i < kWgBox / 16; i += kWgThreads)
                    reinterpret_cast<uint4 *>(smem + s * stage_bytes + b * kWgBox)[i] = make_uint4(0, 0, 0, 0);
            for (int b = nb; b < sh.nb; ++b)
                for (int i = threadIdx.x; i < kWgBox / 16; i += kWgThreads)
                    reinterpret_cast<uint4 *>(smem + s * stage_bytes + (sh.mt * 2 + b) * kWgBox)[i] = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(sh.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(na + nb) * kWgBox;
            for (int blk = blk0; blk < blk1; ++blk) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), tx);
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                for (int b = 0; b < na; ++b) tma_load_2d(sa + b * kWgBox, &map_y, full_bar(stage), m0 + b * 64, blk * kWgNodes);
                for (int b = 0; b < nb; ++b)
                    tma_load_2d(sa + (sh.mt * 2 + b) * kWgBox, &map_x, full_bar(stage), n0 + b * 64, blk * kWgNodes);
                if (++stage == sh.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_mn(sh.nc, BF16), idesc1 = umma_idesc_mn(16, BF16);
            const uint64_t d_ones = umma_desc_mn(smem_u32(s_ones), kWgBox);
            int stage = 0;
            uint32_t phase = 0;
            for (int blk = blk0; blk < blk1; ++blk) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                const uint32_t sb = sa + sh.mt * 2 * kWgBox;
#pragma unroll
                for (int ks = 0; ks < kWgNodes / 16; ++ks) {             // 16 nodes = two 8-row groups = 2048 bytes
                    const uint64_t db = umma_desc_mn(sb + ks * 2048, kWgBox);
                    for (int i = 0; i < sh.mt; ++i) {
                        const uint64_t da = umma_desc_mn(sa + i * 2 * kWgBox + ks * 2048, kWgBox);
                        const uint32_t acc = (blk != blk0 || ks != 0) ? 1u : 0u;
                        tc_mma_f16(tmem_base + (uint32_t)(i * sh.ncp), da, db, idesc, acc);
                        if (bias_here) tc_mma_f16(tmem_base + (uint32_t)(i * sh.ncp + sh.nc), da, d_ones, idesc1, acc);
                    }
                }
                tc_commit(empty_bar(stage));
                if (++stage == sh.stages) { stage = 0; phase ^= 1; }
            }
            tc_commit(done_bar);
        }
    } else {
        // ===== epilogue: TMEM -> this split's fp32 partial [Nout, ldp] (+ bias column ldp - 1 ... see host) =====
        const int q = warp & 3;
        float *out = partial + (size_t)blockIdx.x * Nout * ldp;
        if (blk1 > blk0) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        for (int i = 0; i < sh.mt; ++i) {
            const int row = m0 + i * 128 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * sh.ncp);
            for (int c0 = 0; c0 < sh.nc; c0 += 32) {
                uint32_t v[32];
                if (blk1 > blk0) {
                    tc_ld32(taddr + (uint32_t)c0, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;              // an empty node range contributes zeros
                }
                if (row < Nout) {
                    float *o = out + (size_t)row * ldp + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (c0 + j < sh.nc && n0 + c0 + j < Kin)
                            *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
            if (bias_here) {
                uint32_t v[32];
                if (blk1 > blk0) {
                    tc_ld32(taddr + (uint32_t)sh.nc, v);                 // 16 identical bias columns (+ 16 don't-care)
                    tc_wait_ld();
                } else {
                    v[0] = 0u;
                }
                if (row < Nout) out[(size_t)row * ldp + ldp - 4] = __uint_as_float(v[0]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(sh.tmem_cols) : "memory");
    }
}

// dW[i, j] = Σ_s partial[s][i][j] (split order);  db[i] = Σ_s partial[s][i][ldp - 4]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int Nout, int Kin,
                                                           int ldp, float *__restrict__ dw, int64_t ld_dw,
                                                           float *__restrict__ db) {
    const int64_t n = (int64_t)Nout * (Kin + 1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / (Kin + 1)), j = (int)(t % (Kin + 1));
        if (j == Kin && db == nullptr) continue;
        const float *p = partial + (size_t)i * ldp + (j == Kin ? ldp - 4 : j);
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += p[(size_t)k * Nout * ldp];
        if (j == Kin) db[i] = s;
        else dw[(int64_t)i * ld_dw + j] = s;
    }
}

int wg_box_map(CUtensorMap *map, const void *base, int dtype, int64_t rows, int64_t cols, int64_t ld) {
    return make_map(map, base, dtype, rows, cols, ld, kWgNodes);         // box = 64 columns (128 B) x 64 nodes
}

WgShape wg_shape(int Nout, int Kin, bool bias) {
    WgShape s{};
    const int m_tiles = (Nout + 127) / 128;
    s.mt = std::min(m_tiles, kWgMaxMt);
    s.na = s.mt * 2;
    const int extra = bias ? 16 : 0;
    int nc = (Kin + 15) / 16 * 16;
    nc = std::min(nc, 256);
    nc = std::min(nc, ((512 - 16) / s.mt - extra) / 16 * 16);            // accumulators of all M tiles fit 512 columns
    s.nc = nc;
    s.nb = (nc + 63) / 64;
    s.ncp = nc + extra;
    const int stage_bytes = (s.mt * 2 + s.nb) * kWgBox;
    s.stages = std::max(2, std::min(8, kWgSmemBudget / stage_bytes));
    int cols = 32;
    while (cols < s.mt * s.ncp + 16) cols <<= 1;
    s.tmem_cols = cols;
    return s;
}

int wg_grid(int64_t M, int Nout, int Kin, const WgShape &s, int *n_ranges, int *m_groups, int *per_split) {
    *n_ranges = (Kin + s.nc - 1) / s.nc;
    *m_groups = (Nout + s.mt * 128 - 1) / (s.mt * 128);
    const int64_t node_blocks = (M + kWgNodes - 1) / kWgNodes;
    int splits = (int)std::min<int64_t>(std::max<int64_t>(node_blocks, 1), std::max(1, kNumSMs / (*n_ranges * *m_groups)));
    *per_split = (int)((node_blocks + splits - 1) / splits);
    if (*per_split < 1) *per_split = 1;
    // The tensor core adds into its accumulator with truncation (≈ 2^-26 of the running sum per step, one-sided —
    // measured in gemm_tc.cu): cap the chain at kWgMaxBlocks node blocks (4 steps each => ≤ 3e-5) and let MORE CTAs than
    // SMs run in waves; the split partials are then added in fp32 round-to-nearest by the reduce kernel.
    if (*per_split > kWgMaxBlocks) {                 // whole waves of `splits` CTAs, each at most kWgMaxBlocks long
        const int64_t waves = (node_blocks + (int64_t)splits * kWgMaxBlocks - 1) / ((int64_t)splits * kWgMaxBlocks);
        *per_split = (int)((node_blocks + splits * waves - 1) / (splits * waves));
    }
    splits = (int)std::max<int64_t>(1, (node_blocks + *per_split - 1) / *per_split);
    return splits;
}

}  // namespace
}  // namespace sirgcn

extern "C" size_t sirgcn_gemm_wgrad_workspace_bytes(int64_t m, int32_t n_out, int32_t k_in) {
    using namespace sirgcn;
    if (n_out <= 0 || k_in <= 0 || m < 0) return 0;
    const WgShape s = wg_shape(n_out, k_in, true);
    int nr, mg, per;
    const int splits = wg_grid(m, n_out, k_in, s, &nr, &mg, &per);
    const size_t ldp = ((size_t)k_in + 3) / 4 * 4 + 4;
    return (size_t)splits * n_out * ldp * sizeof(float) + 256;
}

extern "C" int sirgcn_gemm_wgrad_plan(int64_t m, int32_t n_out, int32_t k_in, int32_t with_bias, int32_t *plan) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(plan && m >= 0 && n_out > 0 && k_in > 0, "bad arguments");
    const WgShape s = wg_shape(n_out, k_in, with_bias != 0);
    int nr, mg, per;
    const int splits = wg_grid(m, n_out, k_in, wg_shape(n_out, k_in, true), &nr, &mg, &per);
    nr = (k_in + s.nc - 1) / s.nc;
    const int smem = s.stages * (s.mt * 2 + s.nb) * kWgBox + 2048 + 256 + 1024;
    const int32_t out[10] = {splits, per, nr, mg, s.mt, s.nc, s.ncp, s.stages, s.tmem_cols, smem};
    for (int i = 0; i < 10; ++i) plan[i] = out[i];
    return SIRGCN_OK;
}

extern "C" int sirgcn_gemm_wgrad(const void *dy, int64_t ldy, const void *x, int64_t ldx, int64_t m, int32_t n_out,
                                 int32_t k_in, int32_t dtype, float *dw, int64_t ld_dw, float *db, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(dtype == SIRGCN_BF16 || dtype == SIRGCN_F16, "sirgcn_gemm_wgrad handles bf16/fp16 tables (dtype %d)", dtype);
    SIRGCN_CHECK_ARG(m >= 0 && m < (1LL << 31) && n_out > 0 && k_in > 0, "bad shape m=%lld n_out=%d k_in=%d", (long long)m, n_out, k_in);
    SIRGCN_CHECK_ARG(dw && ld_dw >= k_in, "dW missing / ld too small");
    SIRGCN_CHECK_ARG(n_out % 8 == 0 && k_in % 8 == 0, "n_out and k_in must be multiples of 8 (16-byte rows)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (m == 0) {
        SIRGCN_CUDA(cudaMemset2DAsync(dw, ld_dw * sizeof(float), 0, (size_t)k_in * sizeof(float), n_out, st));
        if (db) SIRGCN_CUDA(cudaMemsetAsync(db, 0, (size_t)n_out * sizeof(float), st));
        return SIRGCN_OK;
    }
    SIRGCN_CHECK_ARG(dy && x && aligned16(dy) && aligned16(x) && ldy % 8 == 0 && ldx % 8 == 0 && ldy >= n_out && ldx >= k_in,
                     "operands must have 16-byte aligned rows");
    const size_t need = sirgcn_gemm_wgrad_workspace_bytes(m, n_out, k_in);
    SIRGCN_CHECK_ARG(workspace && workspace_bytes >= need && aligned16(workspace), "workspace too small: %zu < %zu", workspace_bytes, need);
    const WgShape s = wg_shape(n_out, k_in, db != nullptr);
    int n_ranges, m_groups, per;
    const int splits = wg_grid(m, n_out, k_in, wg_shape(n_out, k_in, true), &n_ranges, &m_groups, &per);
    // the grid is planned with the bias-carrying shape (workspace query); recompute the ranges for the actual one
    n_ranges = (k_in + s.nc - 1) / s.nc;
    CUtensorMap map_y, map_x;
    int rc = wg_box_map(&map_y, dy, dtype, m, n_out, ldy);
    if (rc) return rc;
    rc = wg_box_map(&map_x, x, dtype, m, k_in, ldx);
    if (rc) return rc;
    const int ldp = (k_in + 3) / 4 * 4 + 4;
    const int stage_bytes = (s.mt * 2 + s.nb) * kWgBox;
    const int smem = s.stages * stage_bytes + 2048 + 256 + 1024;
    static std::atomic<int> configured{0};
    if (smem > configured.load(std::memory_order_relaxed)) {
        SIRGCN_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SIRGCN_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.store(smem, std::memory_order_relaxed);
    }
    float *partial = reinterpret_cast<float *>(workspace);
    const dim3 grid((unsigned)splits, (unsigned)n_ranges, (unsigned)m_groups);
    if (dtype == SIRGCN_BF16)
        gemm_wgrad_kernel<true><<<grid, kWgThreads, smem, st>>>(map_y, map_x, partial, (int)m, n_out, k_in, s, db != nullptr, ldp, per);
    else
        gemm_wgrad_kernel<false><<<grid, kWgThreads, smem, st>>>(map_y, map_x, partial, (int)m, n_out, k_in, s, db != nullptr, ldp, per);
    SIRGCN_LAUNCHED();
    const int64_t n = (int64_t)n_out * (k_in + 1);
    wgrad_reduce_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, kNumSMs * 4), 256, 0, st>>>(partial, splits, n_out, k_in, ldp,
                                                                                             dw, ld_dw, db);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}
