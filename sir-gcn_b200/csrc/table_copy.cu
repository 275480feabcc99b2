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

}  // namespace
}  // namespace sirgcn

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
