// Column sums of a feature table in fp32: the bias gradients db_Q = Σ_rows dQ and db_R = Σ_rows dOut of
// nn.Linear's backward (/root/reference/models/conv.py:61,:65 through autograd).  One streaming
// pass at HBM speed; two fixed-order stages (per-CTA partials, then an ordered sum over CTAs) => deterministic.
#include "common.cuh"

namespace sirgcn {
namespace {

constexpr int kColThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kColThreads) colsum_partial_kernel(const T *__restrict__ x, int64_t ld, int64_t m, int n,
                                                                    float *__restrict__ partial) {
    constexpr int NE = VecTraits<T>::N;
    extern __shared__ float sm[];                       // [row lanes][nvec*NE]
    const int nvec = (n + NE - 1) / NE;                 // n is a multiple of NE (16-byte rows)
    const int lanes = kColThreads / nvec;               // row lanes per CTA (nvec <= 256)
    const int vi = threadIdx.x % nvec, rl = threadIdx.x / nvec;
    float acc[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = 0.f;
    if (rl < lanes) {
        const int64_t rows_per_cta = (m + gridDim.x - 1) / gridDim.x;
        const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
        for (int64_t r = r0 + rl; r < r1; r += lanes) {
            float f[NE];
            unpack<T>(ldg_stream(x + r * ld + vi * NE), f);
#pragma unroll
            for (int i = 0; i < NE; ++i) acc[i] += f[i];
        }
#pragma unroll
        for (int i = 0; i < NE; ++i) sm[(rl * nvec + vi) * NE + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nvec * NE; c += kColThreads) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sm[l * nvec * NE + c];
        if (c < n) partial[(int64_t)blockIdx.x * n + c] = s;
    }
}

__global__ void colsum_final_kernel(const float *__restrict__ partial, int parts, int n, float *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s = 0.f;
    for (int p = 0; p < parts; ++p) s += partial[(int64_t)p * n + c];
    out[c] = s;
}

constexpr int kColParts = kNumSMs * 4;

}  // namespace
}  // namespace sirgcn

extern "C" {

size_t sirgcn_colsum_workspace_bytes(int32_t n) { return (size_t)sirgcn::kColParts * (size_t)n * sizeof(float); }

int sirgcn_colsum(const void *x, int64_t ld, int64_t m, int32_t n, int32_t dtype, float *out, void *workspace,
                  size_t workspace_bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(m >= 0 && n > 0 && out, "bad m/n/out");
    const int es = elem_size(dtype), ne = 16 / es;
    SIRGCN_CHECK_ARG(dtype >= SIRGCN_F32 && dtype <= SIRGCN_F16, "bad dtype %d", dtype);
    SIRGCN_CHECK_ARG(n % ne == 0 && n / ne <= kColThreads, "n=%d must be a multiple of %d and at most %d", n, ne, kColThreads * ne);
    SIRGCN_CHECK_ARG(m == 0 || (x && aligned16(x) && (ld * es) % 16 == 0 && ld >= n), "table rows must be 16-byte aligned");
    SIRGCN_CHECK_ARG(workspace && workspace_bytes >= sirgcn_colsum_workspace_bytes(n), "workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int parts = (int)std::min<int64_t>(kColParts, std::max<int64_t>(1, (m + 63) / 64));
    float *partial = reinterpret_cast<float *>(workspace);
    const int nvec = n / ne, lanes = kColThreads / nvec;
    const size_t smem = (size_t)lanes * nvec * ne * sizeof(float);
    switch (dtype) {
        case SIRGCN_F32: colsum_partial_kernel<float><<<parts, kColThreads, smem, st>>>((const float *)x, ld, m, n, partial); break;
        case SIRGCN_BF16: colsum_partial_kernel<__nv_bfloat16><<<parts, kColThreads, smem, st>>>((const __nv_bfloat16 *)x, ld, m, n, partial); break;
        default: colsum_partial_kernel<__half><<<parts, kColThreads, smem, st>>>((const __half *)x, ld, m, n, partial); break;
    }
    SIRGCN_LAUNCHED();
    colsum_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, parts, n, out);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

}  // extern "C"
