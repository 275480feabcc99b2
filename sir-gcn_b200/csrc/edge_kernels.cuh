// Fused SIR-GCN edge stage for sm_100a: gather -> add -> σ -> segment-sum (forward) and the
// recompute backward (dQ over the in-CSR, dK over the out-CSC).  Replaces the DGL
// update_all(UDF message, fn.sum/mean) call of /root/reference/models/conv.py:43-47,:63 and its
// autograd backward.  The |E| x d edge tensor never exists.
//
// Work decomposition
//   * one warp per row; a row's 16-byte feature vectors are spread over G lanes (G = next pow2
//     of the vector count, <= 32) so a warp gathers 32/G neighbour rows per step with 128-bit
//     loads; rows wider than 32 vectors keep VPL vectors per lane.
//   * neighbour ids / edge ids / per-edge coefficients are loaded coalesced (one per lane, 32 per
//     step) and broadcast with shuffles.
//   * U independent gathers are in flight per lane before any arithmetic (latency hiding).
//   * rows longer than `long_threshold` are skipped by the row kernel and processed as
//     fixed-size chunks (one warp per chunk -> fp32 partial) + an ordered finalize, so hubs of a
//     power-law graph are spread over the whole chip; every reduction order is fixed
//     => results are bitwise repeatable, no atomics anywhere.
#pragma once
#include "common.cuh"

namespace sirgcn {
namespace {

enum Mode { kFwd = 0, kBwdQ = 1, kBwdK = 2 };

template <int VPL> struct Unroll { static constexpr int U = VPL == 1 ? 4 : (VPL == 2 ? 2 : 1); };

struct WalkCtx {
    int lane, G, NG, gi, li, nvec;
};

template <typename T, int VPL>
__device__ __forceinline__ void load_row(const void *base, int64_t ld, int64_t row, const WalkCtx &c,
                                         float (&dst)[VPL][VecTraits<T>::N], bool streaming) {
    constexpr int NE = VecTraits<T>::N;
    const T *p = reinterpret_cast<const T *>(base) + row * ld;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int vi = v * c.G + c.li;
        if (vi < c.nvec) {
            uint4 raw = streaming ? ldg_stream(p + vi * NE) : ldg_keep(p + vi * NE);
            unpack<T>(raw, dst[v]);
        } else {
#pragma unroll
            for (int i = 0; i < NE; ++i) dst[v][i] = 0.f;
        }
    }
}

// Accumulates the contributions of positions [beg, end) of one row into acc (per-lane partial:
// lane group gi holds the sum over the edges it visited; caller reduces across groups).
template <typename T, int VPL, int MODE, bool HAS_E>
__device__ __forceinline__ void walk_segment(const sirgcn_edge_args &a, int row, int beg, int end,
                                             const WalkCtx &c, float (&acc)[VPL][VecTraits<T>::N]) {
    constexpr int NE = VecTraits<T>::N;
    constexpr int U = Unroll<VPL>::U;
    constexpr unsigned kFull = 0xffffffffu;

    // row-resident operands
    float self[VPL][NE];
    float ds[VPL][NE];  // backward-dQ: dA[row] * dst_scale[row]
    if (MODE == kBwdK) {
        load_row<T, VPL>(a.k, a.ldk, row, c, self, true);
    } else {
        load_row<T, VPL>(a.q, a.ldq, row, c, self, true);
    }
    if (MODE == kBwdQ) {
        load_row<T, VPL>(a.da, a.lda, row, c, ds, true);
        const float rs = a.dst_scale ? a.dst_scale[row] : 1.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int i = 0; i < NE; ++i) ds[v][i] *= rs;
    }

    const T *tab1 = reinterpret_cast<const T *>(MODE == kBwdK ? a.q : a.k);
    const int64_t ld1 = MODE == kBwdK ? a.ldq : a.ldk;
    const T *tab2 = reinterpret_cast<const T *>(a.da);  // kBwdK only
    const T *etab = HAS_E ? reinterpret_cast<const T *>(a.e) : nullptr;
    T *detab = (HAS_E && MODE == kBwdQ) ? reinterpret_cast<T *>(a.de) : nullptr;
    const float *gscale = MODE == kBwdK ? a.dst_scale : a.src_scale;
    const bool use_eid = HAS_E;

    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_idx = 0, my_eid = 0;
        float my_gs = 1.f;
        if (c.lane < n) {
            my_idx = a.idx[base + c.lane];
            if (use_eid) my_eid = a.eid[base + c.lane];
            if (gscale) my_gs = gscale[my_idx];
        }
        for (int j = 0; j < n; j += c.NG * U) {
            uint4 raw1[U][VPL], raw2[MODE == kBwdK ? U : 1][VPL], rawe[HAS_E ? U : 1][VPL];
            int eids[U];
            float sc[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ej = j + u * c.NG + c.gi;
                const int sl = ej & 31;
                const int node = __shfl_sync(kFull, my_idx, sl);
                sc[u] = __shfl_sync(kFull, my_gs, sl);
                eids[u] = __shfl_sync(kFull, my_eid, sl);
                ok[u] = ej < n;
                if (ok[u]) {
                    const T *p1 = tab1 + (int64_t)node * ld1;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        const int vi = v * c.G + c.li;
                        if (vi < c.nvec) {
                            raw1[u][v] = ldg_stream(p1 + vi * NE);
                            if (MODE == kBwdK) raw2[MODE == kBwdK ? u : 0][v] = ldg_stream(tab2 + (int64_t)node * a.lda + vi * NE);
                            if (HAS_E && etab) rawe[HAS_E ? u : 0][v] = ldg_stream(etab + (int64_t)eids[u] * a.lde + vi * NE);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int vi = v * c.G + c.li;
                    if (vi >= c.nvec) continue;
                    float g1[NE], z[NE];
                    unpack<T>(raw1[u][v], g1);
#pragma unroll
                    for (int i = 0; i < NE; ++i) z[i] = self[v][i] + g1[i];
                    if (HAS_E && etab) {
                        float ev[NE];
                        unpack<T>(rawe[HAS_E ? u : 0][v], ev);
#pragma unroll
                        for (int i = 0; i < NE; ++i) z[i] += ev[i];
                    }
                    if (MODE == kFwd) {
#pragma unroll
                        for (int i = 0; i < NE; ++i) acc[v][i] += sc[u] * act_fwd(z[i], a.act, a.act_param);
                    } else {
                        float up[NE], val[NE];
                        if (MODE == kBwdK) {
                            unpack<T>(raw2[MODE == kBwdK ? u : 0][v], up);
                        } else {
#pragma unroll
                            for (int i = 0; i < NE; ++i) up[i] = ds[v][i];
                        }
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            val[i] = sc[u] * up[i] * act_bwd(z[i], a.act, a.act_param);
                            acc[v][i] += val[i];
                        }
                        if (HAS_E && MODE == kBwdQ && detab) stg_vec(detab + (int64_t)eids[u] * a.ldde + vi * NE, pack<T>(val));
                    }
                }
            }
        }
    }
}

template <int VPL, int NE>
__device__ __forceinline__ void reduce_groups(float (&acc)[VPL][NE], int G) {
    if (VPL == 1) {
        for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
            for (int i = 0; i < NE; ++i) acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], off);
        }
    }
}

__device__ __forceinline__ WalkCtx make_ctx(const sirgcn_edge_args &a, int esize, int vpl) {
    WalkCtx c;
    c.lane = threadIdx.x & 31;
    c.nvec = (a.d * esize + 15) / 16;
    int G = 32;
    if (vpl == 1) {
        G = 1;
        while (G < c.nvec) G <<= 1;
    }
    c.G = G;
    c.NG = 32 / G;
    c.gi = c.lane / G;
    c.li = c.lane % G;
    return c;
}

// ---- one warp per (short) row ---------------------------------------------------------------
template <int VPL> struct MinBlocks { static constexpr int N = VPL == 1 ? 4 : (VPL == 2 ? 2 : 1); };

template <typename T, int VPL, int MODE, bool HAS_E>
__global__ void __launch_bounds__(256, MinBlocks<VPL>::N) edge_rows_kernel(const sirgcn_edge_args a) {
    constexpr int NE = VecTraits<T>::N;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    const int beg = a.indptr[row], end = a.indptr[row + 1];
    if (end - beg > a.long_threshold) return;  // handled by the chunk kernels
    const WalkCtx c = make_ctx(a, sizeof(T), VPL);

    float acc[VPL][NE];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int i = 0; i < NE; ++i) acc[v][i] = 0.f;

    walk_segment<T, VPL, MODE, HAS_E>(a, row, beg, end, c, acc);
    reduce_groups<VPL, NE>(acc, c.G);

    float rs = 1.f;
    if (MODE == kFwd && a.dst_scale) rs = a.dst_scale[row];
    if (MODE == kBwdK && a.src_scale) rs = a.src_scale[row];
    if (c.gi == 0) {
        T *o = reinterpret_cast<T *>(a.out) + (int64_t)row * a.ldo;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int vi = v * c.G + c.li;
            if (vi < c.nvec) {
#pragma unroll
                for (int i = 0; i < NE; ++i) acc[v][i] *= rs;
                stg_vec(o + vi * NE, pack<T>(acc[v]));
            }
        }
    }
}

// ---- one warp per chunk of a long row: fp32 partial ------------------------------------------
template <typename T, int VPL, int MODE, bool HAS_E>
__global__ void __launch_bounds__(256, MinBlocks<VPL>::N) edge_chunks_kernel(const sirgcn_edge_args a) {
    constexpr int NE = VecTraits<T>::N;
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chunk >= a.n_chunks) return;
    const int lrow = a.sched.chunk_lrow[chunk];
    const int row = a.sched.long_rows[lrow];
    const int beg = a.sched.chunk_beg[chunk];
    const int end = min(beg + a.long_threshold, a.indptr[row + 1]);
    const WalkCtx c = make_ctx(a, sizeof(T), VPL);

    float acc[VPL][NE];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int i = 0; i < NE; ++i) acc[v][i] = 0.f;

    walk_segment<T, VPL, MODE, HAS_E>(a, row, beg, end, c, acc);
    reduce_groups<VPL, NE>(acc, c.G);

    if (c.gi == 0) {
        float *o = a.partial + (int64_t)chunk * (c.nvec * NE);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int vi = v * c.G + c.li;
            if (vi < c.nvec) {
#pragma unroll
                for (int i = 0; i < NE; i += 4)
                    *reinterpret_cast<float4 *>(o + vi * NE + i) =
                        make_float4(acc[v][i], acc[v][i + 1], acc[v][i + 2], acc[v][i + 3]);
            }
        }
    }
}

// ---- one CTA per long row: ordered sum of its partials ---------------------------------------
// warp w sums chunks first+w, first+w+8, ... in order; warps are then combined in warp order.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) edge_long_finalize_kernel(const sirgcn_edge_args a) {
    constexpr int NE = VecTraits<T>::N;
    extern __shared__ float smem[];  // [8][nvec*NE]
    const int lrow = blockIdx.x;
    const int row = a.sched.long_rows[lrow];
    const int first = a.sched.long_first[lrow];
    const int nch = a.sched.long_nchunks[lrow];
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const int width = nvec * NE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int col = lane; col < width; col += 32) {
        float s = 0.f;
        for (int ch = warp; ch < nch; ch += 8) s += a.partial[(int64_t)(first + ch) * width + col];
        smem[warp * width + col] = s;
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
        stg_vec(o + vi * NE, pack<T>(r));
    }
}

template <typename T, int VPL, int MODE, bool HAS_E>
int launch_mode(const sirgcn_edge_args &a, cudaStream_t st) {
    constexpr int NE = VecTraits<T>::N;
    if (a.n_rows > 0) {
        const unsigned grid = (unsigned)((a.n_rows + 7) / 8);
        edge_rows_kernel<T, VPL, MODE, HAS_E><<<grid, 256, 0, st>>>(a);
        SIRGCN_LAUNCHED();
    }
    if (a.n_chunks > 0) {
        const unsigned grid = (unsigned)((a.n_chunks + 7) / 8);
        edge_chunks_kernel<T, VPL, MODE, HAS_E><<<grid, 256, 0, st>>>(a);
        SIRGCN_LAUNCHED();
        const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
        const size_t smem = (size_t)8 * nvec * NE * sizeof(float);
        edge_long_finalize_kernel<T, MODE><<<(unsigned)a.n_long, 256, smem, st>>>(a);
        SIRGCN_LAUNCHED();
    }
    return SIRGCN_OK;
}

template <typename T, int MODE>
int launch_vpl(const sirgcn_edge_args &a, cudaStream_t st) {
    const int nvec = (a.d * (int)sizeof(T) + 15) / 16;
    const bool has_e = a.e != nullptr || (MODE == kBwdQ && a.de != nullptr);
#define SIRGCN_LAUNCH_VPL(V) (has_e ? launch_mode<T, V, MODE, true>(a, st) : launch_mode<T, V, MODE, false>(a, st))
    if (nvec <= 32) return SIRGCN_LAUNCH_VPL(1);
    if (nvec <= 64) return SIRGCN_LAUNCH_VPL(2);
    if (nvec <= 128) return SIRGCN_LAUNCH_VPL(4);
#undef SIRGCN_LAUNCH_VPL
    set_error("hidden size %d too wide for the fused edge kernels (max 2048 bytes per row)", a.d);
    return SIRGCN_EUNSUP;
}

}  // namespace

// one translation unit per element type (parallel compilation): edge_f32.cu / edge_bf16.cu / edge_f16.cu
template <typename T>
int edge_launch(const sirgcn_edge_args &a, int mode, cudaStream_t st) {
    switch (mode) {
        case kFwd: return launch_vpl<T, kFwd>(a, st);
        case kBwdQ: return launch_vpl<T, kBwdQ>(a, st);
        default: return launch_vpl<T, kBwdK>(a, st);
    }
}

}  // namespace sirgcn

