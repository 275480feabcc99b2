// Strided row-table copy at HBM speed: dst[r, 0..row_bytes) = src[r, 0..row_bytes) for tables whose rows are whole
// 16-byte vectors but whose row strides differ (one half of a [N, 2·ld] buffer <-> a [N, ld] buffer).  The
// framework's generic strided copy moves such a table at < 2 TB/s; this is 128-bit loads/stores, one vector per
// thread per step, grid sized to the SM count.  Used by the memory-lean backward (function.py) to turn the
// re-made [Q|K] buffer into [dQ|dK] without a second [N, 2·ld] allocation.
#include "common.cuh"

namespace sirgcn {
namespace {

__global__ void __launch_bounds__(256) copy_rows_kernel(char *__restrict__ dst, int64_t dst_pitch,
                                                        const char *__restrict__ src, int64_t src_pitch,
                                                        int vec_per_row, int64_t n_vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const int64_t r = i / vec_per_row;
        const int c = (int)(i - r * vec_per_row);
        stg_vec(dst + r * dst_pitch + c * 16, ldg_stream(src + r * src_pitch + c * 16));
    }
}

// In-place dropout application on a strided row table: t[r, c] = keep[r*d + c] ? t[r, c] * scale : 0, with the
// product taken in fp32 and rounded once (what ATen's fused dropout does: out = src * mask * scale in accscalar_t).
// One thread per 16-byte vector of the table; the keep mask is a dense [rows, d] byte array (torch.bool).
template <typename T>
__global__ void __launch_bounds__(256) mask_scale_kernel(char *__restrict__ tab, int64_t pitch,
                                                         const uint8_t *__restrict__ keep, int d, int vec_per_row,
                                                         int64_t n_vec, float scale) {
    constexpr int NE = VecTraits<T>::N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const int64_t r = i / vec_per_row;
        const int c = (int)(i - r * vec_per_row) * NE;
        char *p = tab + r * pitch + (int64_t)c * sizeof(T);
        float f[NE];
        unpack<T>(*reinterpret_cast<const uint4 *>(p), f);
        const uint8_t *m = keep + r * d + c;
#pragma unroll
        for (int j = 0; j < NE; ++j) f[j] = (c + j < d && m[j]) ? f[j] * scale : 0.f;
        stg_vec(p, pack<T>(f));
    }
}

}  // namespace
}  // namespace sirgcn

extern "C" int sirgcn_mask_scale(void *table, int64_t pitch_bytes, const uint8_t *keep, int64_t rows, int32_t d,
                                 int32_t dtype, float scale, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(rows >= 0 && d > 0, "bad rows/d");
    if (rows == 0) return SIRGCN_OK;
    const int es = elem_size(dtype);
    const int vec_per_row = (d * es + 15) / 16;
    SIRGCN_CHECK_ARG(table && keep && aligned16(table) && pitch_bytes % 16 == 0 && pitch_bytes >= (int64_t)vec_per_row * 16,
                     "table rows must be 16-byte aligned and padded to whole 16-byte vectors");
    const int64_t n_vec = rows * vec_per_row;
    const unsigned grid = (unsigned)std::min<int64_t>((n_vec + 255) / 256, (int64_t)kNumSMs * 32);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *t = reinterpret_cast<char *>(table);
    switch (dtype) {
        case SIRGCN_F32: mask_scale_kernel<float><<<grid, 256, 0, st>>>(t, pitch_bytes, keep, d, vec_per_row, n_vec, scale); break;
        case SIRGCN_BF16: mask_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(t, pitch_bytes, keep, d, vec_per_row, n_vec, scale); break;
        case SIRGCN_F16: mask_scale_kernel<__half><<<grid, 256, 0, st>>>(t, pitch_bytes, keep, d, vec_per_row, n_vec, scale); break;
        default: set_error("bad dtype %d", dtype); return SIRGCN_EINVAL;
    }
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

extern "C" int sirgcn_copy_rows(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                                int64_t row_bytes, int64_t rows, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0, "row_bytes=%lld must be a multiple of 16",
                     (long long)row_bytes);
    if (rows == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(dst && src && aligned16(dst) && aligned16(src) && dst_pitch_bytes % 16 == 0 &&
                         src_pitch_bytes % 16 == 0 && dst_pitch_bytes >= row_bytes && src_pitch_bytes >= row_bytes,
                     "tables must have 16-byte aligned rows and pitches >= row_bytes");
    const int vec_per_row = (int)(row_bytes / 16);
    const int64_t n_vec = rows * vec_per_row;
    const unsigned grid = (unsigned)std::min<int64_t>((n_vec + 255) / 256, (int64_t)kNumSMs * 32);
    copy_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<char *>(dst), dst_pitch_bytes, reinterpret_cast<const char *>(src), src_pitch_bytes,
        vec_per_row, n_vec);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}
