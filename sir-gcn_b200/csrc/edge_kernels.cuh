// Fused SIR-GCN edge stage for sm_100a: gather -> add -> σ -> segment-sum (forward) and the
// recompute backward (dQ over the in-CSR, dK over the out-CSC).  Replaces the DGL
// update_all(UDF message, fn.sum/mean) call of /root/reference/models/conv.py:43-47,:63 and its
// autograd backward.  The |E| x d edge tensor never exists.
//
// Work decomposition (see DESIGN.md "Edge kernels")
//   * rows are grouped at graph-build time into TILES of ~512 work units (one unit per edge + 4 per
//     row), so every warp gets the same amount of work whatever the degree distribution; one warp
//     walks one tile, row after row.  Rows longer than `long_threshold` are skipped there and
//     processed as fixed-size CHUNKS (one warp per chunk -> fp32 partial) + an ordered finalize, so
//     the hubs of a power-law graph are spread over the whole chip.
//   * a row's 16-byte feature vectors are spread over G lanes (G = next pow2 of the vector count,
//     <= 32), so a warp gathers 32/G neighbour rows per step with 128-bit requests.
//   * gathers never touch registers: each lane issues cp.async (LDGSTS, 16 B, L2-only) for exactly
//     the vectors it will later consume into a per-warp ring of S stages in shared memory, so S
//     batches of random rows are in flight per warp while it does the arithmetic of the oldest one —
//     including across row boundaries (the index stream of a tile is contiguous and prefetched two
//     32-entry windows ahead, the q/dA rows of the next row ride in the same ring).
//   * every reduction order is fixed (lane-group shuffle tree, chunk order) => results are bitwise
//     repeatable; there are no atomics anywhere.
#pragma once
#include "common.cuh"

#include <climits>
#include <cstdlib>

namespace sirgcn {
namespace {

enum Mode { kFwd = 0, kBwdQ = 1, kBwdK = 2 };
#ifndef SIRGCN_WARPS
#define SIRGCN_WARPS 4
#endif
#ifndef SIRGCN_STAGES
#define SIRGCN_STAGES 2
#endif
#ifndef SIRGCN_STAGES_Q
#define SIRGCN_STAGES_Q 2
#endif
#ifndef SIRGCN_MIN_CTAS
#define SIRGCN_MIN_CTAS 6
#endif
constexpr int kWarps = SIRGCN_WARPS;  // warps per CTA
constexpr int kTileRowCap = SIRGCN_TILE_WORK / SIRGCN_ROW_COST;     // max rows per tile (graph_build.cu)
constexpr unsigned kFull = 0xffffffffu;

// cp.async with a 32-bit shared address + immediate offset (no generic->shared conversion per copy)
template <int OFF> __device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1], 16;" ::"r"(saddr), "l"(gmem), "n"(OFF) : "memory");
}
template <int OFF> __device__ __forceinline__ void cp_async4(uint32_t saddr, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1], 4;" ::"r"(saddr), "l"(gmem), "n"(OFF) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int OFF> __device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr), "n"(OFF));
    return r;
}
template <int OFF> __device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(r) : "r"(saddr), "n"(OFF));
    return r;
}
template <int OFF> __device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(saddr), "r"(v), "n"(OFF) : "memory");
}
__device__ __forceinline__ int ldg_idx(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- mixed-precision element access: out[i] = c[i] + element i of a 16-byte table vector ---------------
// sm_100a has fp32 += bf16/fp16 adds that read one half of a packed register (SASS FHADD[.BF16] Rn.H0/H1),
// so 16-bit tables are consumed without any unpack instructions and z = q + k is exact in fp32.
template <typename T> __device__ __forceinline__ void add_vec(const uint4 &raw, const float (&c)[VecTraits<T>::N],
                                                              float (&out)[VecTraits<T>::N]);
template <> __device__ __forceinline__ void add_vec<float>(const uint4 &raw, const float (&c)[4], float (&out)[4]) {
    out[0] = c[0] + __uint_as_float(raw.x); out[1] = c[1] + __uint_as_float(raw.y);
    out[2] = c[2] + __uint_as_float(raw.z); out[3] = c[3] + __uint_as_float(raw.w);
}
#define SIRGCN_MIXED_ADD(SUFFIX)                                                                         \
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};                                                  \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                                      \
        unsigned short lo, hi;                                                                           \
        asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w[i]));                                     \
        asm("add.rn.f32." SUFFIX " %0, %1, %2;" : "=f"(out[2 * i]) : "h"(lo), "f"(c[2 * i]));             \
        asm("add.rn.f32." SUFFIX " %0, %1, %2;" : "=f"(out[2 * i + 1]) : "h"(hi), "f"(c[2 * i + 1]));     \
    }
template <> __device__ __forceinline__ void add_vec<__nv_bfloat16>(const uint4 &raw, const float (&c)[8], float (&out)[8]) {
    SIRGCN_MIXED_ADD("bf16")
}
template <> __device__ __forceinline__ void add_vec<__half>(const uint4 &raw, const float (&c)[8], float (&out)[8]) {
    SIRGCN_MIXED_ADD("f16")
}
#undef SIRGCN_MIXED_ADD

// ---- packed 2 x 16-bit helpers of the ReLU fold (see edge_walk_kernel, FOLD) ----------------------------------
// max / compare / add on both halves of a 32-bit register at once (SASS HMNMX2 / HSET2 / HADD2[.BF16_V2])
template <typename T> struct Pk;
template <> struct Pk<float> {      // never instantiated with FOLD; keeps the templates well-formed
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t) { return a; }
    static __device__ __forceinline__ uint32_t gt2(uint32_t a, uint32_t) { return a; }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t) { return a; }
};
template <> struct Pk<__nv_bfloat16> {
    using V = __nv_bfloat162;
    static __device__ __forceinline__ V v(uint32_t x) { return *reinterpret_cast<V *>(&x); }
    static __device__ __forceinline__ uint32_t u(V x) { return *reinterpret_cast<uint32_t *>(&x); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return u(__hmax2(v(a), v(b))); }
    static __device__ __forceinline__ uint32_t gt2(uint32_t a, uint32_t b) { return u(__hgt2(v(a), v(b))); }   // 1.0 / 0.0
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return u(__hadd2(v(a), v(b))); }
};
template <> struct Pk<__half> {
    using V = __half2;
    static __device__ __forceinline__ V v(uint32_t x) { return *reinterpret_cast<V *>(&x); }
    static __device__ __forceinline__ uint32_t u(V x) { return *reinterpret_cast<uint32_t *>(&x); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return u(__hmax2(v(a), v(b))); }
    static __device__ __forceinline__ uint32_t gt2(uint32_t a, uint32_t b) { return u(__hgt2(v(a), v(b))); }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return u(__hadd2(v(a), v(b))); }
};
// largest row / chunk length whose per-element edge count is exact in a 16-bit float (8 / 11 significand bits)
template <typename T> constexpr int fold_max_threshold() { return sizeof(T) == 2 ? 2048 : 0; }
template <> constexpr int fold_max_threshold<__nv_bfloat16>() { return 256; }

// ---- activations as compile-time functors (a runtime switch per element costs issue slots) -------------
template <int ACT> struct Act;
template <> struct Act<SIRGCN_ACT_RELU> {
    static __device__ __forceinline__ float f(float z, float) { return fmaxf(z, 0.f); }
    static __device__ __forceinline__ float d(float z, float) { return z > 0.f ? 1.f : 0.f; }
};
template <> struct Act<SIRGCN_ACT_LEAKY_RELU> {      // also serves Identity (slope 1)
    static __device__ __forceinline__ float f(float z, float p) { return z > 0.f ? z : p * z; }
    static __device__ __forceinline__ float d(float z, float p) { return z > 0.f ? 1.f : p; }
};
template <> struct Act<SIRGCN_ACT_GELU> {
    static __device__ __forceinline__ float f(float z, float) { return act_fwd(z, SIRGCN_ACT_GELU, 0.f); }
    static __device__ __forceinline__ float d(float z, float) { return act_bwd(z, SIRGCN_ACT_GELU, 0.f); }
};

// ---- static configuration of one kernel variant ---------------------------------------------------
// GS = a per-edge scale gathered by neighbour id is present ('sym' in the CSR walks; the CSC walk when dA
// was not pre-scaled by the dQ pass)
template <typename T, int VPL, int MODE, bool HAS_E, bool GS> struct Cfg {
    static constexpr int NE = VecTraits<T>::N;
    static constexpr int NGT = 1 + (MODE == kBwdK ? 1 : 0) + (HAS_E ? 1 : 0);   // gathered tables per edge
    static constexpr int NST = 1 + (MODE == kBwdQ ? 1 : 0);                     // row-resident tables
    static constexpr int U = (4 / (NGT * VPL)) > 0 ? (4 / (NGT * VPL)) : 1;     // edges per lane group per batch
    static constexpr int S = VPL == 4 ? (SIRGCN_STAGES_Q < 3 ? SIRGCN_STAGES_Q : 3)
                                      : (NST == 2 ? SIRGCN_STAGES_Q : SIRGCN_STAGES);   // ring stages
    static constexpr bool DE = HAS_E && MODE == kBwdQ;                          // edge ids kept for the dE store
    static constexpr int kSlot = VPL * 512;                                     // bytes of one row image (32 lanes x 16 B)
    static constexpr int kOffT2 = U * kSlot;                                    // second gathered table (dA, CSC walk)
    static constexpr int kOffTE = (NGT - 1) * U * kSlot;                        // gathered edge term
    static constexpr int kOffSelf = NGT * U * kSlot;
    static constexpr int kOffSc = kOffSelf + NST * kSlot;                       // [U][32] float
    static constexpr int kOffEid = kOffSc + (GS ? U * 128 : 0);                 // [U][32] int
    static constexpr int kStage = kOffEid + (DE ? U * 128 : 0);
};

// per-warp shared memory: [S stages][batch descriptors int2 x cap][row boundaries][row scales]
__host__ __device__ inline int desc_cap(int thr, int B) { return (SIRGCN_TILE_WORK + thr) / B + kTileRowCap + 8; }
__host__ __device__ inline int tail_bytes(int cap) { return (cap * 8 + (kTileRowCap + 4) * 4 + kTileRowCap * 4 + 15) & ~15; }

__host__ __device__ inline int lanes_per_row(int nvec, int vpl) {
    int G = 32;
    if (vpl == 1) {
        G = 1;
        while (G < nvec) G <<= 1;
    }
    return G;
}

// =====================================================================================================
// one warp per tile of short rows / per chunk of a long row
//   pre-pass : the rows of the tile are cut into BATCHES of <= B = NG*U edges of ONE row; the list of
//              batch descriptors {first position, row | count | first | last} is built lane-parallel
//              (warp scan) in shared memory, so the hot loop has no cursor / row-boundary logic.
//   hot loop : S-stage cp.async ring over the batch list; neighbour ids of the next batch are fetched
//              straight into registers one iteration ahead (the lanes of a group read the same word).
// =====================================================================================================
//
// FOLD (ReLU on 16-bit tables, no per-edge scale, no edge term; forward and dQ walks): the issue-bound short-row
// path spends a third of its instructions on z = q + k, σ and the accumulate.  For ReLU they fold algebraically:
//     forward   Σ_v relu(q + k_v) = Σ_v max(k_v, -q) + deg·q     max on PACKED pairs is exact (no rounding), the
//                                                                fp32 accumulate reads the packed halves directly
//     dQ        Σ_v [q + k_v > 0]·g = g ⊙ #{v : k_v > -q}        a packed compare + a packed count (exact up to
//                                                                256 in bf16 / 2048 in fp16 >= the chunk length)
// 1.5 and 1 instructions per element and edge instead of 3; the sign test is the exact one (fl(q+k) > 0 <=> k > -q).
constexpr int kCtrSlots = 4096;
__device__ unsigned int g_unit_ctr[kCtrSlots];                   // work-unit counters of the persistent walks
std::atomic<unsigned int> g_ctr_next{0};

// resident CTAs per SM the register allocation must allow: 8 (64 registers) for the one-vector-per-lane variants
// without an edge term — they fit without spilling and shared memory allows 8 —, SIRGCN_MIN_CTAS (6) for the rest
template <int VPL, bool HAS_E> constexpr int min_ctas() { return (VPL == 1 && !HAS_E) ? 8 : SIRGCN_MIN_CTAS; }

template <typename T, int VPL, int MODE, bool HAS_E, int ACT, bool GS, bool FOLD = false>
__global__ void __launch_bounds__(kWarps * 32, min_ctas<VPL, HAS_E>()) edge_walk_kernel(const sirgcn_edge_args a, const int chunk_mode,
                                                                                       unsigned int *__restrict__ unit_ctr) {
    static_assert(!FOLD || (ACT == SIRGCN_ACT_RELU && !GS && !HAS_E && MODE != kBwdK && sizeof(T) == 2), "FOLD preconditions");
    using C = Cfg<T, VPL, MODE, HAS_E, GS>;
    constexpr int NE = C::NE, U = C::U, S = C::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const int G = lanes_per_row(nvec, VPL), NG = 32 / G, gi = lane / G, li = lane % G;

    const float *gscale = MODE == kBwdK ? a.dst_scale : a.src_scale;             // indexed by idx  (GS)
    const float *rscale = MODE == kBwdK ? a.src_scale : a.dst_scale;             // indexed by row
    const bool has_rs = rscale != nullptr;
    const int Ueff = min(U, 32 / NG);
    const int B = NG * Ueff, logB = 31 - __clz(B);
    const int thr = a.long_threshold;
    const int cap = desc_cap(thr, B);

    // table-gradient mode (dQ walk, edge term from a small table): per-warp fp32 accumulators [type][VPL][lane][NE],
    // every lane owns its 16-byte... NE-float slot => no conflicts, no atomics, fixed order
    const int n_et = (C::DE && a.de == nullptr && a.de_partial != nullptr) ? a.n_etypes : 0;
    const int tacc_bytes = n_et * VPL * 32 * NE * (int)sizeof(float);
    unsigned char *wsm = smem_raw + (size_t)warp * (S * C::kStage + tail_bytes(cap) + tacc_bytes);
    int2 *s_desc = reinterpret_cast<int2 *>(wsm + S * C::kStage);
    int *s_ptr = reinterpret_cast<int *>(s_desc + cap);
    float *s_rs = reinterpret_cast<float *>(s_ptr + kTileRowCap + 4);
    float *s_tacc = reinterpret_cast<float *>(wsm + S * C::kStage + tail_bytes(cap)) + lane * NE;   // this lane's slot of type 0, vector 0
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(wsm) + lane * 16;   // this lane's 16-B column of the ring
    const uint32_t ring4 = (uint32_t)__cvta_generic_to_shared(wsm) + lane * 4;   // this lane's 4-B column

    // ---- which rows: PERSISTENT warps — every warp owns its shared-memory region and never synchronises with the
    // others, so each warp loops over work units on its own, taking the next unit number from a device counter
    // (units are handed out in index order, like CTAs of a plain grid: same L2 locality, no tail imbalance; the
    // counter's latency hides under the unit being walked).  A grid of exactly the resident CTAs removes the CTA
    // launch / drain gaps of one-CTA-per-4-tiles grids (ncu r01: 36 % achieved occupancy of 50 % theoretical).
    // unit_ctr == nullptr: plain grid, one unit per warp. ---------------------------------------------------------
    const int n_units = chunk_mode ? a.n_chunks : a.n_tiles;
    int unit = 0, r_beg = 0, nrows = 0, total = 0;
    const bool write_scaled = MODE == kBwdQ && a.da_scaled != nullptr && !chunk_mode;   // long rows: finalize kernel
    auto setup = [&]() -> bool {
        if (chunk_mode) {
            r_beg = a.sched.long_rows[a.sched.chunk_lrow[unit]];
            nrows = 1;
            const int beg = a.sched.chunk_beg[unit];
            if (lane == 0) {
                s_ptr[0] = beg;
                s_ptr[1] = min(beg + thr, a.indptr[r_beg + 1]);
            }
        } else {
            r_beg = a.tile_row[unit];
            nrows = a.tile_row[unit + 1] - r_beg;
            if (nrows <= 0) return false;
            for (int i = lane; i <= nrows; i += 32) s_ptr[i] = a.indptr[r_beg + i];
        }
        __syncwarp();

        // ---- pre-pass: batch list -----------------------------------------------------------------------
        total = 0;
        for (int j0 = 0; j0 < nrows; j0 += 32) {
            const int i = j0 + lane;
            int beg = 0, deg = 0, nb = 0;
            if (i < nrows) {
                beg = s_ptr[i];
                deg = s_ptr[i + 1] - beg;
                if (chunk_mode || deg <= thr) nb = max(1, (deg + B - 1) >> logB);    // empty row: one batch of 0 edges (A[u] = 0)
                if (has_rs) s_rs[i] = rscale[r_beg + i];
            }
            int incl = nb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const int off = total + incl - nb;
            for (int k = 0; k < nb; ++k)
                s_desc[off + k] = make_int2(beg + (k << logB),
                                            i | (min(B, deg - (k << logB)) << 8) | (k == 0 ? 1 << 16 : 0) | (k == nb - 1 ? 1 << 17 : 0));
            total += __shfl_sync(kFull, incl, 31);
        }
        __syncwarp();
        return true;
    };

    // ---- lane constants --------------------------------------------------------------------------------------
    bool vok[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) vok[v] = v * G + li < nvec;
    int slot[U];          // position of this lane group's u-th edge inside a batch (64 => never valid)
#pragma unroll
    for (int u = 0; u < U; ++u) slot[u] = (u < Ueff && vok[0]) ? gi * Ueff + u : 64;   // a group's edges are consecutive
    const int esz = (int)sizeof(T);
    const char *tab1 = reinterpret_cast<const char *>(MODE == kBwdK ? a.q : a.k) + li * 16;   // gathered by idx
    const uint32_t ld1 = (uint32_t)((MODE == kBwdK ? a.ldq : a.ldk) * esz);
    const char *tab2 = reinterpret_cast<const char *>(a.da) + li * 16;                        // CSC walk: gathered by idx
    const uint32_t ld2 = (uint32_t)(a.lda * esz);
    const char *selft = reinterpret_cast<const char *>(MODE == kBwdK ? a.k : a.q) + li * 16;  // row resident
    const uint32_t ldself = (uint32_t)((MODE == kBwdK ? a.ldk : a.ldq) * esz);
    const char *etab = HAS_E ? reinterpret_cast<const char *>(a.e) + li * 16 : nullptr;
    const uint32_t lde = (uint32_t)(a.lde * esz);
    char *detab = (C::DE && a.de) ? reinterpret_cast<char *>(a.de) + li * 16 : nullptr;   // NULL: no dE rows wanted
    const int32_t *idxp = a.idx + gi * Ueff;
    const int32_t *eidp = HAS_E ? a.eid + gi * Ueff : nullptr;
    const int vstride = G * 16;                                  // bytes between a lane's consecutive vectors of one row
    // keep the 64-bit bases in registers (otherwise they are re-derived from the constant bank per access)
    asm("" : "+l"(tab1)); asm("" : "+l"(tab2)); asm("" : "+l"(selft)); asm("" : "+l"(idxp));
    if (HAS_E) { asm("" : "+l"(etab)); asm("" : "+l"(eidp)); }

    // ---- neighbour ids of the next batch to issue, one iteration ahead ---------------------------------------
    int nd[U], ne[HAS_E ? U : 1];
    auto prefetch = [&](int b) {
        if (b >= total) return;
        const int2 d = s_desc[b];
        const int cnt = (d.y >> 8) & 255;
        const int32_t *ip = idxp + d.x;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] < cnt) {
                nd[u] = ldg_idx(ip + u);
                if (HAS_E) ne[HAS_E ? u : 0] = ldg_idx(eidp + d.x + u);
            }
        }
    };

    int istage = 0, cstage = 0;                                  // ring slots of the next issue / consume
    auto issue = [&](int b) {
        const int2 d = s_desc[b];
        const int cnt = (d.y >> 8) & 255;
        const uint32_t st = ring + istage * C::kStage;
        const uint32_t st4 = ring4 + istage * C::kStage;
        istage = istage + 1 == S ? 0 : istage + 1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] < cnt) {
                const int node = nd[u];
                const char *p1 = tab1 + (uint64_t)(uint32_t)node * ld1;
                const char *p2 = MODE == kBwdK ? tab2 + (uint64_t)(uint32_t)node * ld2 : nullptr;
                const char *pe = (HAS_E && etab) ? etab + (uint64_t)(uint32_t)ne[HAS_E ? u : 0] * lde : nullptr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if (v > 0 && !vok[v]) continue;
                    switch (u) {   // immediate ring offsets
#define SIRGCN_CP(UU)                                                                                      \
    case UU:                                                                                               \
        cp_async16<(UU * VPL) * 512>(st + v * 512, p1 + v * vstride);                                      \
        if (MODE == kBwdK) cp_async16<C::kOffT2 + (UU * VPL) * 512>(st + v * 512, p2 + v * vstride);       \
        if (HAS_E && etab) cp_async16<C::kOffTE + (UU * VPL) * 512>(st + v * 512, pe + v * vstride);       \
        break;
                        SIRGCN_CP(0) SIRGCN_CP(1) SIRGCN_CP(2) SIRGCN_CP(3)
#undef SIRGCN_CP
                    }
                }
                if (GS) {
                    const float *gp = gscale + node;
                    switch (u) {
                        case 0: cp_async4<C::kOffSc>(st4, gp); break;
                        case 1: cp_async4<C::kOffSc + 128>(st4, gp); break;
                        case 2: cp_async4<C::kOffSc + 256>(st4, gp); break;
                        default: cp_async4<C::kOffSc + 384>(st4, gp); break;
                    }
                }
                if (C::DE) {
                    switch (u) {
                        case 0: sts32<C::kOffEid>(st4, ne[0]); break;
                        case 1: sts32<C::kOffEid + 128>(st4, ne[HAS_E && U > 1 ? 1 : 0]); break;
                        case 2: sts32<C::kOffEid + 256>(st4, ne[HAS_E && U > 2 ? 2 : 0]); break;
                        default: sts32<C::kOffEid + 384>(st4, ne[HAS_E && U > 3 ? 3 : 0]); break;
                    }
                }
            }
        }
        if (d.y & (1 << 16)) {                                   // first batch of a row: its own operands
            const uint32_t row = (uint32_t)(r_beg + (d.y & 255));
            const char *ps = selft + (uint64_t)row * ldself;
            const char *pa = MODE == kBwdQ ? tab2 + (uint64_t)row * ld2 : nullptr;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (!vok[v]) continue;
                cp_async16<C::kOffSelf>(st + v * 512, ps + v * vstride);
                if (MODE == kBwdQ) cp_async16<C::kOffSelf + C::kSlot>(st + v * 512, pa + v * vstride);
            }
        }
        prefetch(b + 1);
    };

    // ---- arithmetic of one landed batch -------------------------------------------------------------------
    float self[FOLD ? 1 : VPL][NE], ds[MODE == kBwdQ ? VPL : 1][NE], acc[VPL][NE];
    uint4 negq[FOLD ? VPL : 1], cntp[FOLD && MODE == kBwdQ ? VPL : 1];       // FOLD: packed -q, packed edge counts
    int rowcnt = 0;                                                           // FOLD forward: edges of the row so far
    float rs = 1.f;
    const float ap = a.act_param;
    const float zero[NE] = {};
    auto consume = [&](int b) {
        const int2 d = s_desc[b];
        const int cnt = (d.y >> 8) & 255;
        const uint32_t st = ring + cstage * C::kStage;
        const uint32_t st4 = ring4 + cstage * C::kStage;
        cstage = cstage + 1 == S ? 0 : cstage + 1;
        if (d.y & (1 << 16)) {
            if (has_rs) rs = s_rs[d.y & 255];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (vok[v]) {
                    if constexpr (FOLD) {
                        const uint4 rq = lds128<C::kOffSelf>(st + v * 512);
                        negq[v] = make_uint4(rq.x ^ 0x80008000u, rq.y ^ 0x80008000u, rq.z ^ 0x80008000u, rq.w ^ 0x80008000u);
                        if (MODE == kBwdQ) cntp[MODE == kBwdQ ? v : 0] = make_uint4(0u, 0u, 0u, 0u);
                    } else {
                        add_vec<T>(lds128<C::kOffSelf>(st + v * 512), zero, self[FOLD ? 0 : v]);
                    }
                    if (MODE == kBwdQ) {
                        add_vec<T>(lds128<C::kOffSelf + C::kSlot>(st + v * 512), zero, ds[MODE == kBwdQ ? v : 0]);
#pragma unroll
                        for (int i = 0; i < NE; ++i) ds[MODE == kBwdQ ? v : 0][i] *= rs;
                        if (write_scaled && gi == 0) {
                            char *o = reinterpret_cast<char *>(a.da_scaled) + (uint64_t)(uint32_t)(r_beg + (d.y & 255)) * (uint32_t)(a.ldds * esz) + li * 16;
                            stg_vec(o + v * vstride, pack<T>(ds[MODE == kBwdQ ? v : 0]));
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NE; ++i) acc[v][i] = 0.f;
            }
            rowcnt = 0;
        }
        if constexpr (FOLD) rowcnt += cnt;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] >= cnt) continue;
            float sc = 1.f;
            if (GS) sc = __uint_as_float(u == 0 ? lds32<C::kOffSc>(st4) : u == 1 ? lds32<C::kOffSc + 128>(st4)
                                         : u == 2 ? lds32<C::kOffSc + 256>(st4) : lds32<C::kOffSc + 384>(st4));
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (v > 0 && !vok[v]) continue;
                uint4 r1, r2 = make_uint4(0, 0, 0, 0), re = make_uint4(0, 0, 0, 0);
                switch (u) {
#define SIRGCN_LD(UU)                                                                               \
    case UU:                                                                                        \
        r1 = lds128<(UU * VPL) * 512>(st + v * 512);                                                \
        if (MODE == kBwdK) r2 = lds128<C::kOffT2 + (UU * VPL) * 512>(st + v * 512);                 \
        if (HAS_E) re = lds128<C::kOffTE + (UU * VPL) * 512>(st + v * 512);                         \
        break;
                    SIRGCN_LD(0) SIRGCN_LD(1) SIRGCN_LD(2) default: SIRGCN_LD(3)
#undef SIRGCN_LD
                }
                if constexpr (FOLD) {
                    const uint4 nq = negq[v];
                    if (MODE == kFwd) {
                        const uint4 m = make_uint4(Pk<T>::max2(r1.x, nq.x), Pk<T>::max2(r1.y, nq.y),
                                                   Pk<T>::max2(r1.z, nq.z), Pk<T>::max2(r1.w, nq.w));
                        add_vec<T>(m, acc[v], acc[v]);
                    } else {
                        uint4 &c = cntp[MODE == kBwdQ ? v : 0];
                        c.x = Pk<T>::add2(c.x, Pk<T>::gt2(r1.x, nq.x));
                        c.y = Pk<T>::add2(c.y, Pk<T>::gt2(r1.y, nq.y));
                        c.z = Pk<T>::add2(c.z, Pk<T>::gt2(r1.z, nq.z));
                        c.w = Pk<T>::add2(c.w, Pk<T>::gt2(r1.w, nq.w));
                    }
                    continue;
                }
                float z[NE];
                add_vec<T>(r1, self[FOLD ? 0 : v], z);
                if (HAS_E && a.e) add_vec<T>(re, z, z);
                if (MODE == kFwd) {
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        if (GS) acc[v][i] += sc * Act<ACT>::f(z[i], ap);
                        else acc[v][i] += Act<ACT>::f(z[i], ap);
                    }
                } else if (ACT == SIRGCN_ACT_RELU && !C::DE) {
                    // σ' ∈ {0,1}: predicated accumulate of the upstream gradient, no multiply
                    float t[NE];
                    if (MODE == kBwdK) {
                        if (GS) {
                            add_vec<T>(r2, zero, t);
#pragma unroll
                            for (int i = 0; i < NE; ++i) t[i] = fmaf(sc, t[i], acc[v][i]);
                        } else {
                            add_vec<T>(r2, acc[v], t);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NE; ++i) t[i] = GS ? fmaf(sc, ds[MODE == kBwdQ ? v : 0][i], acc[v][i]) : acc[v][i] + ds[MODE == kBwdQ ? v : 0][i];
                    }
#pragma unroll
                    for (int i = 0; i < NE; ++i) acc[v][i] = z[i] > 0.f ? t[i] : acc[v][i];
                } else {
                    float up[NE], val[NE];
                    if (MODE == kBwdK) {
                        add_vec<T>(r2, zero, up);
                    } else {
#pragma unroll
                        for (int i = 0; i < NE; ++i) up[i] = ds[MODE == kBwdQ ? v : 0][i];
                    }
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        val[i] = (GS ? sc * up[i] : up[i]) * Act<ACT>::d(z[i], ap);
                        acc[v][i] += val[i];
                    }
                    if (C::DE) {
                        if (detab || n_et) {
                            const int eidv = (int)(u == 0 ? lds32<C::kOffEid>(st4) : u == 1 ? lds32<C::kOffEid + 128>(st4)
                                                   : u == 2 ? lds32<C::kOffEid + 256>(st4) : lds32<C::kOffEid + 384>(st4));
                            if (detab) {
                                stg_vec(detab + (uint64_t)(uint32_t)eidv * (uint32_t)(a.ldde * esz) + v * vstride, pack<T>(val));
                            } else {                             // eidv = the edge's TYPE: add into this lane's slot
                                float *t = s_tacc + (eidv * VPL + v) * (32 * NE);
#pragma unroll
                                for (int i = 0; i < NE; i += 4) {
                                    float4 c = *reinterpret_cast<float4 *>(t + i);
                                    c.x += val[i]; c.y += val[i + 1]; c.z += val[i + 2]; c.w += val[i + 3];
                                    *reinterpret_cast<float4 *>(t + i) = c;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (d.y & (1 << 17)) {                                   // last batch of a row: reduce + store
            if constexpr (FOLD && MODE == kBwdQ) {
                if (VPL == 1) {                                  // the lane groups' counts add up exactly (<= row length)
                    for (int off = 16; off >= G; off >>= 1) {
                        uint4 &c = cntp[0];
                        c.x = Pk<T>::add2(c.x, __shfl_xor_sync(kFull, c.x, off));
                        c.y = Pk<T>::add2(c.y, __shfl_xor_sync(kFull, c.y, off));
                        c.z = Pk<T>::add2(c.z, __shfl_xor_sync(kFull, c.z, off));
                        c.w = Pk<T>::add2(c.w, __shfl_xor_sync(kFull, c.w, off));
                    }
                }
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    float cf[NE];
                    add_vec<T>(cntp[MODE == kBwdQ ? v : 0], zero, cf);
#pragma unroll
                    for (int i = 0; i < NE; ++i) acc[v][i] = ds[MODE == kBwdQ ? v : 0][i] * cf[i];
                }
            } else {
                if (VPL == 1) {
                    for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
                        for (int i = 0; i < NE; ++i) acc[0][i] += __shfl_xor_sync(kFull, acc[0][i], off);
                    }
                }
                if constexpr (FOLD) {                            // + deg·q  (−deg·(−q): the packed negation is at hand)
                    const float c = -(float)rowcnt;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        float nq[NE];
                        add_vec<T>(negq[v], zero, nq);
#pragma unroll
                        for (int i = 0; i < NE; ++i) acc[v][i] = fmaf(c, nq[i], acc[v][i]);
                    }
                }
            }
            if (gi == 0) {
                if (chunk_mode) {
                    float *o = a.partial + (int64_t)unit * (nvec * NE) + li * NE;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (!vok[v]) continue;
#pragma unroll
                        for (int i = 0; i < NE; i += 4)
                            *reinterpret_cast<float4 *>(o + v * G * NE + i) =
                                make_float4(acc[v][i], acc[v][i + 1], acc[v][i + 2], acc[v][i + 3]);
                    }
                } else {
                    const float os = MODE == kBwdQ ? 1.f : rs;   // bwd-dQ folded dst_scale into ds
                    char *o = reinterpret_cast<char *>(a.out) + (uint64_t)(uint32_t)(r_beg + (d.y & 255)) * (uint32_t)(a.ldo * esz) + li * 16;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (!vok[v]) continue;
#pragma unroll
                        for (int i = 0; i < NE; ++i) acc[v][i] *= os;
                        if (a.accumulate)      // phased walks over source blocks (partition.py): out += this block
                            add_vec<T>(*reinterpret_cast<const uint4 *>(o + v * vstride), acc[v], acc[v]);
                        stg_vec(o + v * vstride, pack<T>(acc[v]));
                    }
                }
            }
        }
    };

    // ---- S-stage software pipeline; one commit per iteration keeps wait_group<S-1> exact ---------------------
    unsigned int ticket = 0;                                     // lane 0: the next unit of this warp
    if (unit_ctr) {
        if (lane == 0) ticket = atomicAdd(unit_ctr, 1u);
        unit = (int)__shfl_sync(kFull, ticket, 0);
    } else {
        unit = blockIdx.x * kWarps + warp;
    }
#pragma unroll 1
    for (; unit < n_units; unit = unit_ctr ? (int)__shfl_sync(kFull, ticket, 0) : n_units) {
        if (unit_ctr && lane == 0) ticket = atomicAdd(unit_ctr, 1u);       // consumed at the end of this unit
        if (!setup()) continue;
        istage = cstage = 0;
        if (C::DE && n_et) {                                     // zero this lane's slots
            for (int j = 0; j < n_et * VPL; ++j)
#pragma unroll
                for (int i = 0; i < NE; i += 4) *reinterpret_cast<float4 *>(s_tacc + j * (32 * NE) + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        prefetch(0);
#pragma unroll 1
        for (int it = 0; it < total + S - 1; ++it) {
            if (it < total) issue(it);
            cp_async_commit();
            if (it >= S - 1) {
                cp_async_wait<S - 1>();
                consume(it - (S - 1));
            }
        }
        __syncwarp();                                            // the next unit's pre-pass rewrites the batch list
        if (C::DE && n_et) {
            // this unit's sums per edge type: the lane groups are combined in group order by the lanes of group 0
            if (gi == 0 && vok[0]) {
                const int width = nvec * NE;
                float *o = a.de_partial + ((int64_t)(unit + (chunk_mode ? a.n_tiles : 0)) * n_et) * width + li * NE;
                for (int t = 0; t < n_et; ++t) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (!vok[v]) continue;
                        float r[NE];
#pragma unroll
                        for (int i = 0; i < NE; ++i) r[i] = 0.f;
                        const float *src = s_tacc + (t * VPL + v) * (32 * NE);
                        for (int gg = 0; gg < NG; ++gg)
#pragma unroll
                            for (int i = 0; i < NE; ++i) r[i] += src[gg * G * NE + i];
#pragma unroll
                        for (int i = 0; i < NE; i += 4)
                            *reinterpret_cast<float4 *>(o + (int64_t)t * width + v * G * NE + i) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---- long rows: ordered sum of the chunk partials -------------------------------------------------
// One WARP per long row (a power-law graph has tens of thousands of rows just above the threshold with a
// handful of chunks each): lane l owns 16-byte vectors l, l+32, ...; chunks are summed in chunk order with
// four interleaved accumulators (ch mod 4), combined in a fixed order => bitwise repeatable.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) edge_long_finalize_kernel(const sirgcn_edge_args a) {
    constexpr int NE = VecTraits<T>::N;
    const int lrow = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (lrow >= a.n_long) return;
    const int lane = threadIdx.x & 31;
    const int row = a.sched.long_rows[lrow];
    const int first = a.sched.long_first[lrow];
    const int nch = a.sched.long_nchunks[lrow];
    if (nch > SIRGCN_BIG_CHUNKS) return;                 // hubs: edge_big_finalize_kernel
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const int width = nvec * NE;
    float rs = 1.f;
    if (MODE == kFwd && a.dst_scale) rs = a.dst_scale[row];
    if (MODE == kBwdK && a.src_scale) rs = a.src_scale[row];
    T *o = reinterpret_cast<T *>(a.out) + (int64_t)row * a.ldo;
    for (int vi = lane; vi < nvec; vi += 32) {
        float s[4][NE];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < NE; ++i) s[j][i] = 0.f;
        const float *p = a.partial + (int64_t)first * width + vi * NE;
        int ch = 0;
        for (; ch + 4 <= nch; ch += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < NE; i += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(p + (int64_t)(ch + j) * width + i);
                    s[j][i] += t.x; s[j][i + 1] += t.y; s[j][i + 2] += t.z; s[j][i + 3] += t.w;
                }
        }
        for (int j = 0; ch < nch; ++ch, ++j) {
#pragma unroll
            for (int i = 0; i < NE; i += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(p + (int64_t)ch * width + i);
                s[j][i] += t.x; s[j][i + 1] += t.y; s[j][i + 2] += t.z; s[j][i + 3] += t.w;
            }
        }
        float r[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) r[i] = ((s[0][i] + s[1][i]) + (s[2][i] + s[3][i])) * rs;
        if (a.accumulate) add_vec<T>(*reinterpret_cast<const uint4 *>(o + vi * NE), r, r);
        stg_vec(o + vi * NE, pack<T>(r));
        if (MODE == kBwdQ && a.da_scaled) {      // dA[row] * dst_scale[row] for the CSC walk (may alias dA: all chunks are done)
            const float ds = a.dst_scale ? a.dst_scale[row] : 1.f;
            float g[NE];
            unpack<T>(ldg_keep(reinterpret_cast<const T *>(a.da) + (int64_t)row * a.lda + vi * NE), g);
#pragma unroll
            for (int i = 0; i < NE; ++i) g[i] *= ds;
            stg_vec(reinterpret_cast<T *>(a.da_scaled) + (int64_t)row * a.ldds + vi * NE, pack<T>(g));
        }
    }
}

// Hubs (more than SIRGCN_BIG_CHUNKS chunks): one CTA per row, taken from the schedule's big list by a
// fixed-size grid.  Warp w sums chunks w, w+8, ... with four interleaved accumulators over float4 columns;
// the 8 warps are combined in warp order through shared memory => fixed order, bitwise repeatable.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) edge_big_finalize_kernel(const sirgcn_edge_args a) {
    constexpr int NE = VecTraits<T>::N;
    extern __shared__ float smem[];                      // [8][width]
    const int n_big = a.sched.big_lrows[0];
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const int width = nvec * NE, w4 = width / 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x; b < n_big; b += gridDim.x) {
        const int lrow = a.sched.big_lrows[1 + b];
        const int row = a.sched.long_rows[lrow];
        const int first = a.sched.long_first[lrow];
        const int nch = a.sched.long_nchunks[lrow];
        const float4 *p = reinterpret_cast<const float4 *>(a.partial + (int64_t)first * width);
        for (int c4 = lane; c4 < w4; c4 += 32) {
            float4 s[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            int ch = warp;
            for (; ch + 24 < nch; ch += 32) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t = p[(int64_t)(ch + 8 * j) * w4 + c4];
                    s[j].x += t.x; s[j].y += t.y; s[j].z += t.z; s[j].w += t.w;
                }
            }
            for (int j = 0; ch < nch; ch += 8, ++j) {
                const float4 t = p[(int64_t)ch * w4 + c4];
                s[j].x += t.x; s[j].y += t.y; s[j].z += t.z; s[j].w += t.w;
            }
            float4 r;
            r.x = (s[0].x + s[1].x) + (s[2].x + s[3].x); r.y = (s[0].y + s[1].y) + (s[2].y + s[3].y);
            r.z = (s[0].z + s[1].z) + (s[2].z + s[3].z); r.w = (s[0].w + s[1].w) + (s[2].w + s[3].w);
            reinterpret_cast<float4 *>(smem + warp * width)[c4] = r;
        }
        __syncthreads();
        float rs = 1.f;
        if (MODE == kFwd && a.dst_scale) rs = a.dst_scale[row];
        if (MODE == kBwdK && a.src_scale) rs = a.src_scale[row];
        T *o = reinterpret_cast<T *>(a.out) + (int64_t)row * a.ldo;
        for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
            float r[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += smem[w * width + vi * NE + i];
                r[i] = s * rs;
            }
            if (a.accumulate) add_vec<T>(*reinterpret_cast<const uint4 *>(o + vi * NE), r, r);
            stg_vec(o + vi * NE, pack<T>(r));
            if (MODE == kBwdQ && a.da_scaled) {
                const float ds = a.dst_scale ? a.dst_scale[row] : 1.f;
                float g[NE];
                unpack<T>(ldg_keep(reinterpret_cast<const T *>(a.da) + (int64_t)row * a.lda + vi * NE), g);
#pragma unroll
                for (int i = 0; i < NE; ++i) g[i] *= ds;
                stg_vec(reinterpret_cast<T *>(a.da_scaled) + (int64_t)row * a.ldds + vi * NE, pack<T>(g));
            }
        }
        __syncthreads();
    }
}

template <typename T, int VPL, int MODE, bool HAS_E, int ACT, bool GS, bool FOLD = false>
int launch_gs(const sirgcn_edge_args &a, cudaStream_t st) {
    using C = Cfg<T, VPL, MODE, HAS_E, GS>;
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const int NG = 32 / lanes_per_row(nvec, VPL);
    const int B = NG * std::min(C::U, 32 / NG);
    static const size_t smem_pad = getenv("SIRGCN_SMEM_PAD") ? (size_t)atoi(getenv("SIRGCN_SMEM_PAD")) : 0;   // occupancy experiments
    const int n_et = (C::DE && a.de == nullptr && a.de_partial != nullptr) ? a.n_etypes : 0;
    const size_t smem = (size_t)kWarps * (C::S * C::kStage + tail_bytes(desc_cap(a.long_threshold, B)) +
                                          (size_t)n_et * VPL * 32 * VecTraits<T>::N * sizeof(float)) + smem_pad;
    if (smem > 227 * 1024) {
        set_error("long_threshold %d needs %zu bytes of shared memory per CTA (max 232448)", a.long_threshold, smem);
        return SIRGCN_EUNSUP;
    }
    auto kern = edge_walk_kernel<T, VPL, MODE, HAS_E, ACT, GS, FOLD>;
    static std::atomic<int> configured{0};           // per instantiation: raise the opt-in limit only when needed
    if ((int)smem > configured.load(std::memory_order_relaxed)) {   // (one process drives one GPU)
        SIRGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store((int)smem, std::memory_order_relaxed);
    }
    // persistent grid: as many CTAs as are resident at once (occupancy query, cached per instantiation and smem size)
    static std::atomic<int> resident{0}, resident_smem{-1};
    if (resident_smem.load(std::memory_order_relaxed) != (int)smem) {
        int per_sm = 0;
        SIRGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem));
        resident.store(std::max(1, per_sm) * kNumSMs, std::memory_order_relaxed);
        resident_smem.store((int)smem, std::memory_order_relaxed);
    }
    static const bool no_persist = getenv("SIRGCN_NO_PERSIST") != nullptr;       // A/B measurements
    const int cap = resident.load(std::memory_order_relaxed);
    // unit counters: a ring of words in static device memory (the library allocates nothing); a launch zeroes its
    // word on its own stream just before it runs, and a word comes round again only after kCtrSlots later launches
    static unsigned int *ctr_bases[64] = {};                     // per device: a __device__ symbol has one address per GPU
    int dev = 0;
    SIRGCN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!ctr_bases[dev]) SIRGCN_CUDA(cudaGetSymbolAddress(reinterpret_cast<void **>(&ctr_bases[dev]), g_unit_ctr));
    unsigned int *const ctr_base = ctr_bases[dev];
    auto walk = [&](int n_units, int chunk_mode) -> int {
        const int grid_all = (n_units + kWarps - 1) / kWarps;
        if (no_persist || (a.flags & SIRGCN_WALK_PLAIN_GRID) || grid_all <= cap) {   // (or: everything is resident at once anyway)
            kern<<<(unsigned)grid_all, kWarps * 32, smem, st>>>(a, chunk_mode, nullptr);
        } else {
            unsigned int *ctr = ctr_base + (g_ctr_next.fetch_add(1, std::memory_order_relaxed) % kCtrSlots);
            SIRGCN_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
            kern<<<(unsigned)cap, kWarps * 32, smem, st>>>(a, chunk_mode, ctr);
        }
        SIRGCN_LAUNCHED();
        return SIRGCN_OK;
    };
    if (a.n_tiles > 0) {
        const int rc = walk(a.n_tiles, 0);
        if (rc) return rc;
    }
    if (a.n_chunks > 0) {
        const int rc = walk(a.n_chunks, 1);
        if (rc) return rc;
        edge_long_finalize_kernel<T, MODE><<<(unsigned)((a.n_long + 7) / 8), 256, 0, st>>>(a);
        SIRGCN_LAUNCHED();
        const int big_grid = std::min(a.n_chunks / SIRGCN_BIG_CHUNKS, kNumSMs * 4);   // n_big <= n_chunks / 33
        if (big_grid > 0) {
            edge_big_finalize_kernel<T, MODE><<<big_grid, 256, (size_t)8 * nvec * VecTraits<T>::N * sizeof(float), st>>>(a);
            SIRGCN_LAUNCHED();
        }
    }
    return SIRGCN_OK;
}

template <typename T, int VPL, int MODE, bool HAS_E, int ACT>
int launch_act(const sirgcn_edge_args &a, cudaStream_t st) {
    const float *gscale = MODE == kBwdK ? a.dst_scale : a.src_scale;
    if constexpr (ACT == SIRGCN_ACT_RELU && !HAS_E && sizeof(T) == 2 && MODE != kBwdK) {
        static const bool no_fold = getenv("SIRGCN_NO_FOLD") != nullptr;         // A/B measurements
        if (!gscale && !no_fold && a.long_threshold <= fold_max_threshold<T>())
            return launch_gs<T, VPL, MODE, HAS_E, ACT, false, true>(a, st);
    }
    return gscale ? launch_gs<T, VPL, MODE, HAS_E, ACT, true>(a, st) : launch_gs<T, VPL, MODE, HAS_E, ACT, false>(a, st);
}

template <typename T, int VPL, int MODE, bool HAS_E>
int launch_mode(const sirgcn_edge_args &a, cudaStream_t st) {
    switch (a.act) {
        case SIRGCN_ACT_RELU: return launch_act<T, VPL, MODE, HAS_E, SIRGCN_ACT_RELU>(a, st);
        case SIRGCN_ACT_GELU: return launch_act<T, VPL, MODE, HAS_E, SIRGCN_ACT_GELU>(a, st);
        case SIRGCN_ACT_IDENTITY: {                      // identity == leaky ReLU with slope 1 (exactly)
            sirgcn_edge_args b = a;
            b.act_param = 1.f;
            return launch_act<T, VPL, MODE, HAS_E, SIRGCN_ACT_LEAKY_RELU>(b, st);
        }
        default: return launch_act<T, VPL, MODE, HAS_E, SIRGCN_ACT_LEAKY_RELU>(a, st);
    }
}

template <typename T, int MODE>
int launch_vpl(const sirgcn_edge_args &a, cudaStream_t st) {
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const bool has_e = a.e != nullptr || (MODE == kBwdQ && (a.de != nullptr || (a.n_etypes > 0 && a.de_partial != nullptr)));
#define SIRGCN_LAUNCH_VPL(V) (has_e ? launch_mode<T, V, MODE, true>(a, st) : launch_mode<T, V, MODE, false>(a, st))
    if (nvec <= 32) return SIRGCN_LAUNCH_VPL(1);
    if (nvec <= 64) return SIRGCN_LAUNCH_VPL(2);
    if (nvec <= 128) return SIRGCN_LAUNCH_VPL(4);
#undef SIRGCN_LAUNCH_VPL
    set_error("hidden size %d too wide for the fused edge kernels (max 2048 bytes per row)", a.d);
    return SIRGCN_EUNSUP;
}

}  // namespace

// one translation unit per (element type, mode): edge_<type>_<mode>.cu  (parallel compilation)
template <typename T, int MODE> int edge_launch(const sirgcn_edge_args &a, cudaStream_t st) {
    return launch_vpl<T, MODE>(a, st);
}

}  // namespace sirgcn
