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
#include <cuda.h>

#include "common.cuh"

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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes, 8-row groups of 1024 B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
    return d;
}

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

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap *map, const void *base, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_rows,
             bool store = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is unavailable in this driver");
        return SIRGCN_EUNSUP;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == SIRGCN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, store ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a %lld x %lld table, ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return SIRGCN_EINVAL;
    }
    return SIRGCN_OK;
}

}  // namespace
}  // namespace sirgcn

extern "C" int sirgcn_gemm_tn(const void *a, int64_t lda, const void *b, int64_t ldb, void *c, int64_t ldc,
                              const float *bias, int64_t m, int32_t n, int32_t k, int32_t dtype, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(dtype == SIRGCN_BF16 || dtype == SIRGCN_F16, "sirgcn_gemm_tn handles bf16/fp16 tables (dtype %d)", dtype);
    SIRGCN_CHECK_ARG(m >= 0 && m < (1LL << 31) && n > 0 && k > 0, "bad shape m=%lld n=%d k=%d", (long long)m, n, k);
    if (m == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(a && b && c, "a/b/c is NULL");
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
