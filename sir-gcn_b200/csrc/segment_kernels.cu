// Split ("generic") edge path: gather-add -> arbitrary callable (PyTorch) -> segment reduce.
// Used when σ is not one of the elementwise activations the fused kernels know
// (e.g. Sequential(ReLU, Linear, ReLU), /root/reference/synthetic-datasets/dictionary-lookup/model.py:17),
// for agg_type 'max'/'min' (/root/reference/models/conv.py:47: W_R is applied per edge) and for
// SIRConvBase/SIREConvBase (conv.py:137-221).  Edge tensors are materialised here, in CSR position
// order; no alignment or padding requirements (scalar element accesses, lanes over columns).
#include "common.cuh"

namespace sirgcn {
namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

constexpr int kWarpsPerBlock = 8;

// z[p] = a[asel[p]] (+ b[bsel[p]]) (+ c[csel[p]])   — one warp per position
template <typename T>
__global__ void __launch_bounds__(256) gather_add_kernel(int64_t num_pos, const int32_t *__restrict__ asel,
                                                         const int32_t *__restrict__ bsel, const int32_t *__restrict__ csel,
                                                         const T *__restrict__ a, int64_t lda, const T *__restrict__ b, int64_t ldb,
                                                         const T *__restrict__ c, int64_t ldc, T *__restrict__ z, int64_t ldz, int d) {
    const int64_t p = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= num_pos) return;
    const int lane = threadIdx.x & 31;
    const T *ra = a + (int64_t)asel[p] * lda;
    const T *rb = b ? b + (int64_t)bsel[p] * ldb : nullptr;
    const T *rc = c ? c + (int64_t)csel[p] * ldc : nullptr;
    T *rz = z + p * ldz;
    for (int col = lane; col < d; col += 32) {
        float v = to_f(ra[col]);
        if (rb) v += to_f(rb[col]);
        if (rc) v += to_f(rc[col]);
        rz[col] = from_f<T>(v);
    }
}

// out[u] = ds[u] * sum_{p in row u} ss[idx[p]] * m[perm ? perm[p] : p]   — one warp per row, ordered
template <typename T>
__global__ void __launch_bounds__(256) segment_sum_kernel(int n_rows, const int32_t *__restrict__ indptr,
                                                          const int32_t *__restrict__ idx, const int32_t *__restrict__ perm,
                                                          const T *__restrict__ m, int64_t ldm, T *__restrict__ out, int64_t ldo,
                                                          int d, const float *__restrict__ ds, const float *__restrict__ ss) {
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int beg = indptr[row], end = indptr[row + 1];
    const float rs = ds ? ds[row] : 1.f;
    for (int col0 = 0; col0 < d; col0 += 128) {      // 4 columns per lane per sweep
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = beg; p < end; ++p) {
            const int64_t pos = perm ? perm[p] : p;
            const float s = ss ? ss[idx[p]] : 1.f;
            const T *rm = m + pos * ldm;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int col = col0 + t * 32 + lane;
                if (col < d) acc[t] += s * to_f(rm[col]);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int col = col0 + t * 32 + lane;
            if (col < d) out[(int64_t)row * ldo + col] = from_f<T>(acc[t] * rs);
        }
    }
}

// first strict extremum in CSR position order wins (DGL's CPU SpMMCmpCsr order); empty rows -> 0 / -1
template <typename T, bool IS_MIN>
__global__ void __launch_bounds__(256) segment_minmax_kernel(int n_rows, const int32_t *__restrict__ indptr,
                                                             const T *__restrict__ m, int64_t ldm, T *__restrict__ out,
                                                             int64_t ldo, int32_t *__restrict__ arg, int64_t ldarg, int d) {
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int beg = indptr[row], end = indptr[row + 1];
    for (int col = lane; col < d; col += 32) {
        float best = 0.f;
        int where = -1;
        for (int p = beg; p < end; ++p) {
            const float v = to_f(m[(int64_t)p * ldm + col]);
            const bool better = where < 0 || (IS_MIN ? v < best : v > best);
            if (better) { best = v; where = p; }
        }
        out[(int64_t)row * ldo + col] = from_f<T>(best);
        arg[(int64_t)row * ldarg + col] = where;
    }
}

// dm[p, c] = arg[rsel[p], c] == p ? dout[rsel[p], c] : 0   — one warp per position
template <typename T>
__global__ void __launch_bounds__(256) segment_minmax_bwd_kernel(int64_t num_pos, const int32_t *__restrict__ rsel,
                                                                 const T *__restrict__ dout, int64_t lddo,
                                                                 const int32_t *__restrict__ arg, int64_t ldarg,
                                                                 T *__restrict__ dm, int64_t lddm, int d) {
    const int64_t p = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= num_pos) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = rsel[p];
    for (int col = lane; col < d; col += 32) {
        const bool win = arg[row * ldarg + col] == (int32_t)p;
        dm[p * lddm + col] = win ? dout[row * lddo + col] : from_f<T>(0.f);
    }
}

inline unsigned warp_grid(int64_t n) { return (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock); }

#define SIRGCN_DISPATCH_DTYPE(dtype, ...)                          \
    switch (dtype) {                                               \
        case SIRGCN_F32: { using T = float; __VA_ARGS__; break; }  \
        case SIRGCN_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
        case SIRGCN_F16: { using T = __half; __VA_ARGS__; break; } \
        default: set_error("bad dtype %d", dtype); return SIRGCN_EINVAL; \
    }

}  // namespace
}  // namespace sirgcn

extern "C" {

int sirgcn_gather_add(int64_t num_pos, const int32_t *asel, const int32_t *bsel, const int32_t *csel,
                      const void *a, int64_t lda, const void *b, int64_t ldb, const void *c, int64_t ldc,
                      void *z, int64_t ldz, int32_t d, int32_t dtype, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_pos >= 0 && d > 0, "bad num_pos/d");
    if (num_pos == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(a && asel && z, "a/asel/z is NULL");
    SIRGCN_CHECK_ARG((b == nullptr) || bsel, "b given without bsel");
    SIRGCN_CHECK_ARG((c == nullptr) || csel, "c given without csel");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SIRGCN_DISPATCH_DTYPE(dtype, gather_add_kernel<T><<<warp_grid(num_pos), 256, 0, st>>>(
        num_pos, asel, bsel, csel, (const T *)a, lda, (const T *)b, ldb, (const T *)c, ldc, (T *)z, ldz, d));
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_segment_sum(int32_t n_rows, const int32_t *indptr, const int32_t *idx, const int32_t *perm,
                       const void *m, int64_t ldm, void *out, int64_t ldo, int32_t d, int32_t dtype,
                       const float *dst_scale, const float *src_scale, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(n_rows >= 0 && d > 0, "bad n_rows/d");
    if (n_rows == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(indptr && out, "indptr/out is NULL");
    SIRGCN_CHECK_ARG(!src_scale || idx, "src_scale given without idx");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SIRGCN_DISPATCH_DTYPE(dtype, segment_sum_kernel<T><<<warp_grid(n_rows), 256, 0, st>>>(
        n_rows, indptr, idx, perm, (const T *)m, ldm, (T *)out, ldo, d, dst_scale, src_scale));
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_segment_minmax(int32_t n_rows, const int32_t *indptr, const void *m, int64_t ldm,
                          void *out, int64_t ldo, int32_t *arg, int64_t ldarg, int32_t d, int32_t dtype,
                          int32_t is_min, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(n_rows >= 0 && d > 0, "bad n_rows/d");
    if (n_rows == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(indptr && out && arg, "indptr/out/arg is NULL");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_min) {
        SIRGCN_DISPATCH_DTYPE(dtype, segment_minmax_kernel<T, true><<<warp_grid(n_rows), 256, 0, st>>>(
            n_rows, indptr, (const T *)m, ldm, (T *)out, ldo, arg, ldarg, d));
    } else {
        SIRGCN_DISPATCH_DTYPE(dtype, segment_minmax_kernel<T, false><<<warp_grid(n_rows), 256, 0, st>>>(
            n_rows, indptr, (const T *)m, ldm, (T *)out, ldo, arg, ldarg, d));
    }
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_segment_minmax_bwd(int64_t num_pos, const int32_t *rsel, const void *dout, int64_t lddo,
                              const int32_t *arg, int64_t ldarg, void *dm, int64_t lddm,
                              int32_t d, int32_t dtype, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_pos >= 0 && d > 0, "bad num_pos/d");
    if (num_pos == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(rsel && dout && arg && dm, "NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SIRGCN_DISPATCH_DTYPE(dtype, segment_minmax_bwd_kernel<T><<<warp_grid(num_pos), 256, 0, st>>>(
        num_pos, rsel, (const T *)dout, lddo, arg, ldarg, (T *)dm, lddm, d));
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

}  // extern "C"
