// Dense projections of the SIR-GCN layer on the 5th-generation tensor cores (sm_100a):
//     C[M, N] = A[M, K] · B[N, K]^T (+ bias[N]),   A/B/C bf16 or fp16, fp32 accumulation in TMEM.
// Replaces the cuBLAS calls behind nn.Linear for linear_query ‖ linear_key (one concatenated [W_Q;W_K]
// projection), linear_relation and their input gradients (/root/reference/models/conv.py:60-61,:65;
// SURVEY.md K1, K2, K9 and the dgrad half of K12).  Both operands are K-major, so the same kernel serves
//     forward  [Q|K] = H · [W_Q;W_K]^T + [b_Q|0]         (B = the weight as stored, [out, in])
//     dgrad    dH    = dY · W = dY · (W^T)^T              (B = W^T, a tiny host-side transpose)
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles of A (128 x 64) and B (BN x 64) into a 3-stage
//            128B-swizzled shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1   allocates TMEM, issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN<=256, K=16) from one
//            elected lane; tcgen05.commit releases ring slots and publishes finished accumulators
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> +bias -> bf16/fp16 -> 128B-swizzled
//            staging tile in shared memory -> cp.async.bulk.tensor store (coalesced, clips the M/N tails);
//            two TMEM accumulator stages let tile i+1's MMAs overlap tile i's epilogue
// These GEMMs are skinny (K = d_in <= 512, N = 2d or d_out): they sit at the memory/compute ridge
// (DESIGN.md §2.4), so the tile is chosen to read A once and write C once; B is L2-resident.
#include "tc_common.cuh"

namespace sirgcn {
namespace {

constexpr int kBM = 128;          // rows of C per tile = UMMA_M
constexpr int kBK = 64;           // K elements per ring stage (= one 128-byte swizzle row of 16-bit elements)
constexpr int kStages = 3;
constexpr int kMaxBN = 256;
constexpr int kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kBBytes = kMaxBN * kBK * 2;     // 32 KB (box may be smaller)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 192;
constexpr int kCBoxBytes = kBM * 64 * 2;      // one 128 x 64 output box (128-byte swizzled rows) = 16 KB
constexpr int kCBytes = (kMaxBN / 64) * kCBoxBytes;   // output staging for the TMA store: 64 KB
constexpr int kSmemBytes = kStages * kStageBytes + kCBytes + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment slack*/;

// instruction descriptor for kind::f16: D = fp32, A/B = bf16 or fp16, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t umma_idesc(int bn, bool bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                                    // D format: F32
    d |= (bf16 ? 1u : 0u) << 7;                      // A format
    d |= (bf16 ? 1u : 0u) << 10;                     // B format
    d |= (uint32_t)(bn >> 3) << 17;                  // N >> 3
    d |= (uint32_t)(kBM >> 4) << 24;                 // M >> 4
    return d;
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const float *__restrict__ bias, int M, int N, int K, int bn) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *s_c = smem + kStages * kStageBytes;                   // 1024-byte aligned output staging
    float *s_bias = reinterpret_cast<float *>(s_c + kCBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_c + kCBytes + 1024);
    // bars: full[0..S), empty[S..2S), tmem_full[2S..2S+2), tmem_empty[2S+2..2S+4), then the TMEM base word
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStages + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + bn - 1) / bn;
    const int tiles = m_tiles * n_tiles, kblocks = (K + kBK - 1) / kBK;
    const int bnp = (bn + 31) & ~31;                  // accumulator stage stride: the epilogue reads 32-column groups
    int tmem_cols = 32;
    while (tmem_cols < 2 * bnp) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1) {   // TMEM allocation: one full warp, address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
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
            const uint32_t tx = (uint32_t)(kABytes + bn * kBK * 2);
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * bn;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), tx);
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    tma_load_2d(sa, &map_a, full_bar(stage), kb * kBK, m0);
                    tma_load_2d(sa + kABytes, &map_b, full_bar(stage), kb * kBK, n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(bn, BF16);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(as), aphase ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * bnp);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);          // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = umma_desc(sa), db = umma_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)          // 32 bytes along K inside the swizzle row
                        tc_mma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(empty_bar(stage));                // slot free when these MMAs have read it
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(as));                       // accumulator complete
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes [32q, 32q+32), q = warp % 4 =====
        const int q = warp & 3;
        int as = 0;
        uint32_t aphase = 0;
        int bias_n0 = -1;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * bn;
            if (bias != nullptr && n0 != bias_n0) {             // stage this column block's bias once
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int j = threadIdx.x - 64; j < bnp; j += 128) s_bias[j] = (j < bn && n0 + j < N) ? bias[n0 + j] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_n0 = n0;
            }
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            // the previous tile's TMA store must have finished reading the staging tile
            if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int r = q * 32 + lane;                        // row inside the tile = TMEM lane
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * bnp);
            const uint32_t srow = smem_u32(s_c) + (uint32_t)r * 128u;
            for (int c0 = 0; c0 < bnp; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + (uint32_t)c0, v);
                tc_wait_ld();
                const uint32_t box = srow + (uint32_t)(c0 >> 6) * kCBoxBytes;
                const int chunk0 = (c0 & 63) >> 3;              // 16-byte chunk inside the 128-byte row
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * j + i]) + (bias ? s_bias[c0 + 8 * j + i] : 0.f);
                    const uint32_t dst = box + ((uint32_t)((chunk0 + j) ^ (r & 7)) << 4);     // SWIZZLE_128B
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2<BF16>(f[0], f[1])),
                                 "r"(pack2<BF16>(f[2], f[3])), "r"(pack2<BF16>(f[4], f[5])), "r"(pack2<BF16>(f[6], f[7])) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));         // TMEM stage drained: the next MMAs may overwrite it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA engine
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                for (int b = 0; b * 64 < bn; ++b)
                    tma_store_2d(&map_c, smem_u32(s_c) + (uint32_t)b * kCBoxBytes, n0 + b * 64, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output tiles written
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}


// =====================================================================================================================
// fp32 tables: C[M, N] = A[M, K] · B[N, K]^T (+ bias) in fp32 on the tensor cores by the 3xTF32 split
//     a ≈ a_hi + a_lo,  a_hi = a rounded to TF32 (10 explicit mantissa bits), a_lo = (a - a_hi) rounded to TF32
//     a·b ≈ a_hi·b_hi + a_lo·b_hi + a_hi·b_lo + a_lo·b_lo          (representation error 2^-23 per operand)
// fp32 accumulation in TMEM; measured ~1e-7 relative against fp64, inside the 1e-5 fp32 parity target that plain TF32
// (1e-3) misses and that made round 1 keep the library SGEMM (tcgen05 has no IEEE fp32 MMA).  The reference computes
// these projections with cuBLAS SGEMM (/root/reference/models/conv.py:60-61,:65).
// Same structure as gemm_tn_kernel plus a SPLITTER warpgroup: TMA lands the raw fp32 tiles (128-byte swizzled rows of
// 32 floats); warps 6-9 rewrite each tile in place as its hi part and write the lo part to a twin tile at the same
// offsets (an elementwise rewrite keeps the swizzle), fence to the async proxy and hand the stage to the MMA warp,
// which issues four kind::tf32 MMAs per 8-wide K step.
// =====================================================================================================================
constexpr int kF32BK = 32;                          // fp32 elements per 128-byte swizzle row
constexpr int kF32MaxBN = 128;
constexpr int kF32Stages = 2;
constexpr int kF32Tile = 128 * 128;                 // one operand tile: 128 rows x 128 bytes = 16 KB
constexpr int kF32StageBytes = 4 * kF32Tile;        // A_hi, A_lo, B_hi, B_lo
constexpr int kF32CBytes = (kF32MaxBN / 32) * kF32Tile;     // output staging: 4 boxes of 128 x 32 floats
constexpr int kF32Threads = 320;
constexpr int kF32SmemBytes = kF32Stages * kF32StageBytes + kF32CBytes + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*slack*/;

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int bn) {
    uint32_t d = 0;
    d |= 1u << 4;                                    // D format: F32
    d |= 2u << 7;                                    // A format: TF32
    d |= 2u << 10;                                   // B format: TF32
    d |= (uint32_t)(bn >> 3) << 17;                  // N >> 3
    d |= (uint32_t)(kBM >> 4) << 24;                 // M >> 4
    return d;
}

// hi = x rounded to the nearest TF32 (10 explicit mantissa bits), lo = (x - hi) rounded to the nearest TF32 as well:
// the tensor core TRUNCATES the low 13 mantissa bits of what it reads, so a residual left at 13 significant bits would
// lose up to 2^-21·|x|; rounded here, x = hi + lo to within 2^-23·|x|
__device__ __forceinline__ float rn_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(x - hi);
}

__global__ void __launch_bounds__(kF32Threads, 1)
gemm_tn_f32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_c, const float *__restrict__ bias, int M, int N, int K, int bn) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *s_c = smem + kF32Stages * kF32StageBytes;
    float *s_bias = reinterpret_cast<float *>(s_c + kF32CBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_c + kF32CBytes + 1024);
    // bars: full[S], split[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base word
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * kF32Stages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto split_bar = [&](int s) { return bar0 + 8u * (kF32Stages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * kF32Stages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * kF32Stages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * kF32Stages + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + bn - 1) / bn;
    const int tiles = m_tiles * n_tiles, kblocks = (K + kF32BK - 1) / kF32BK;
    const int bnp = (bn + 31) & ~31;
    // two accumulators per stage: the hi·hi products, and the three small correction products.  The tensor core adds
    // into its fp32 accumulator with TRUNCATION (measured: a one-sided error growing linearly with the number of
    // accumulation steps, 9e-7 at K = 128 with all four products in one accumulator); the corrections are 2^-11 of
    // the result, so their accumulator's truncation is negligible, and the main one sees a quarter of the steps.  The
    // epilogue adds the two in fp32 (round to nearest).
    int tmem_cols = 32;
    while (tmem_cols < 4 * bnp) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kF32Stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(split_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b_bytes = (uint32_t)bn * 128u;

    if (warp == 0) {
        // ===== TMA producer: raw fp32 tiles into the A_hi / B_hi slots =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * bn;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), (uint32_t)kF32Tile + b_bytes);
                    const uint32_t sa = smem_u32(smem + stage * kF32StageBytes);
                    tma_load_2d(sa, &map_a, full_bar(stage), kb * kF32BK, m0);
                    tma_load_2d(sa + 2 * kF32Tile, &map_b, full_bar(stage), kb * kF32BK, n0);
                    if (++stage == kF32Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: three TF32 products per K step =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(bn);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * 2 * bnp), tmem_c = tmem_d + (uint32_t)bnp;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(split_bar(stage), phase);         // hi / lo tiles are in place and visible to the async proxy
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kF32StageBytes);
                    const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + kF32Tile);
                    const uint64_t b_hi = umma_desc(sa + 2 * kF32Tile), b_lo = umma_desc(sa + 3 * kF32Tile);
#pragma unroll
                    for (int k = 0; k < kF32BK / 8; ++k) {      // 8 floats = 32 bytes along K inside the swizzle row
                        const uint64_t o = (uint64_t)(k * 2);
                        tc_mma_tf32(tmem_c, a_lo + o, b_lo + o, idesc, (kb | k) != 0);     // 2^-22: smallest first
                        tc_mma_tf32(tmem_c, a_lo + o, b_hi + o, idesc, 1);
                        tc_mma_tf32(tmem_c, a_hi + o, b_lo + o, idesc, 1);
                        tc_mma_tf32(tmem_d, a_hi + o, b_hi + o, idesc, (kb | k) != 0);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == kF32Stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(as));
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 6) {
        // ===== splitter: tile -> (hi in place, lo in the twin tile) =====
        const int tid = threadIdx.x - 192;
        int stage = 0;
        uint32_t phase = 0;
        const int b_vec = (int)(b_bytes >> 4);
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t sa = smem_u32(smem + stage * kF32StageBytes);
                for (int part = 0; part < 2; ++part) {
                    const uint32_t base = sa + (uint32_t)part * 2u * kF32Tile;
                    const int n_vec = part == 0 ? kF32Tile / 16 : b_vec;
                    for (int i = tid; i < n_vec; i += 128) {
                        const uint32_t addr = base + (uint32_t)i * 16u;
                        float4 v, hi, lo;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
                        split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr + kF32Tile), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> visible to the tensor core
                mbar_arrive(split_bar(stage));
                if (++stage == kF32Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes [32q, 32q+32), q = warp % 4 =====
        const int q = warp & 3;
        int as = 0;
        uint32_t aphase = 0;
        int bias_n0 = -1;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * bn;
            if (bias != nullptr && n0 != bias_n0) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int j = threadIdx.x - 64; j < bnp; j += 128) s_bias[j] = (j < bn && n0 + j < N) ? bias[n0 + j] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_n0 = n0;
            }
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int r = q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * bnp);
            const uint32_t srow = smem_u32(s_c) + (uint32_t)r * 128u;
            for (int c0 = 0; c0 < bnp; c0 += 32) {              // one 128 x 32 box per 32 accumulator columns
                uint32_t v[32], vc[32];
                tc_ld32(taddr + (uint32_t)c0, v);
                tc_ld32(taddr + (uint32_t)(bnp + c0), vc);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vc[i]));
                const uint32_t box = srow + (uint32_t)(c0 >> 5) * kF32Tile;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[i] = __uint_as_float(v[4 * j + i]) + (bias ? s_bias[c0 + 4 * j + i] : 0.f);
                    const uint32_t dst = box + ((uint32_t)(j ^ (r & 7)) << 4);        // SWIZZLE_128B
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                for (int b = 0; b * 32 < bn; ++b)
                    tma_store_2d(&map_c, smem_u32(s_c) + (uint32_t)b * kF32Tile, n0 + b * 32, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace
}  // namespace sirgcn

extern "C" int sirgcn_gemm_tn(const void *a, int64_t lda, const void *b, int64_t ldb, void *c, int64_t ldc,
                              const float *bias, int64_t m, int32_t n, int32_t k, int32_t dtype, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(dtype == SIRGCN_BF16 || dtype == SIRGCN_F16 || dtype == SIRGCN_F32, "bad dtype %d", dtype);
    SIRGCN_CHECK_ARG(m >= 0 && m < (1LL << 31) && n > 0 && k > 0, "bad shape m=%lld n=%d k=%d", (long long)m, n, k);
    if (m == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(a && b && c, "a/b/c is NULL");
    if (dtype == SIRGCN_F32) {
        SIRGCN_CHECK_ARG(n % 4 == 0 && k % 4 == 0, "n and k must be multiples of 4 (16-byte rows): n=%d k=%d", n, k);
        SIRGCN_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(c) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 &&
                             lda >= k && ldb >= k && ldc >= n, "operands must have 16-byte aligned rows");
        const int bn = n >= kF32MaxBN ? kF32MaxBN : (n + 15) / 16 * 16;
        CUtensorMap map_a, map_b, map_c;
        int rc = make_map(&map_a, a, dtype, m, k, lda, kBM);
        if (rc) return rc;
        rc = make_map(&map_b, b, dtype, n, k, ldb, bn);
        if (rc) return rc;
        rc = make_map(&map_c, c, dtype, m, n, ldc, kBM, true);
        if (rc) return rc;
        const int tiles = (int)((m + kBM - 1) / kBM) * ((n + bn - 1) / bn);
        static std::atomic<bool> configured32{false};
        if (!configured32.load(std::memory_order_relaxed)) {
            SIRGCN_CUDA(cudaFuncSetAttribute(gemm_tn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF32SmemBytes));
            configured32.store(true, std::memory_order_relaxed);
        }
        gemm_tn_f32_kernel<<<std::min(tiles, kNumSMs), kF32Threads, kF32SmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
            map_a, map_b, map_c, bias, (int)m, n, k, bn);
        SIRGCN_LAUNCHED();
        return SIRGCN_OK;
    }
    SIRGCN_CHECK_ARG(n % 8 == 0 && k % 8 == 0, "n and k must be multiples of 8 (16-byte rows): n=%d k=%d", n, k);
    SIRGCN_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(c) && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 &&
                         lda >= k && ldb >= k && ldc >= n, "operands must have 16-byte aligned rows");
    int bn = n >= kMaxBN ? kMaxBN : (n + 15) / 16 * 16;
    CUtensorMap map_a, map_b, map_c;
    int rc = make_map(&map_a, a, dtype, m, k, lda, kBM);
    if (rc) return rc;
    rc = make_map(&map_b, b, dtype, n, k, ldb, bn);
    if (rc) return rc;
    rc = make_map(&map_c, c, dtype, m, n, ldc, kBM, true);
    if (rc) return rc;
    const int tiles = (int)((m + kBM - 1) / kBM) * ((n + bn - 1) / bn);
    const int grid = std::min(tiles, kNumSMs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static std::atomic<bool> configured{false};
    if (!configured.load(std::memory_order_relaxed)) {
        SIRGCN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        SIRGCN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured.store(true, std::memory_order_relaxed);
    }
    if (dtype == SIRGCN_BF16) {
        gemm_tn_kernel<true><<<grid, kGemmThreads, kSmemBytes, st>>>(map_a, map_b, map_c, bias, (int)m, n, k, bn);
    } else {
        gemm_tn_kernel<false><<<grid, kGemmThreads, kSmemBytes, st>>>(map_a, map_b, map_c, bias, (int)m, n, k, bn);
    }
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}
