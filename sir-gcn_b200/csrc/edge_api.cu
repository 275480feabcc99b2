// C-ABI entry points of the fused edge stage (argument validation + dtype dispatch).
#include "common.cuh"

namespace sirgcn {
enum Mode { kFwd = 0, kBwdQ = 1, kBwdK = 2 };
template <typename T, int MODE> int edge_launch(const sirgcn_edge_args &a, cudaStream_t st);
#define SIRGCN_DECL(T)                                                                  \
    extern template int edge_launch<T, kFwd>(const sirgcn_edge_args &, cudaStream_t);   \
    extern template int edge_launch<T, kBwdQ>(const sirgcn_edge_args &, cudaStream_t);  \
    extern template int edge_launch<T, kBwdK>(const sirgcn_edge_args &, cudaStream_t);
SIRGCN_DECL(float)
SIRGCN_DECL(__nv_bfloat16)
SIRGCN_DECL(__half)
#undef SIRGCN_DECL

namespace {

int validate(const sirgcn_edge_args *a, int mode) {
    SIRGCN_CHECK_ARG(a != nullptr, "args is NULL");
    SIRGCN_CHECK_ARG(a->n_rows >= 0 && a->d > 0, "bad n_rows=%d / d=%d", a->n_rows, a->d);
    SIRGCN_CHECK_ARG(a->dtype >= SIRGCN_F32 && a->dtype <= SIRGCN_F16, "bad dtype %d", a->dtype);
    SIRGCN_CHECK_ARG(a->act >= SIRGCN_ACT_IDENTITY && a->act <= SIRGCN_ACT_GELU, "bad act %d", a->act);
    SIRGCN_CHECK_ARG(a->long_threshold >= 32, "long_threshold must be >= 32");
    if (a->n_rows == 0) return SIRGCN_OK;
    const int es = elem_size(a->dtype);
    const int ldmin = ((a->d * es + 15) / 16) * 16 / es;
    SIRGCN_CHECK_ARG(a->indptr && a->idx, "indptr/idx is NULL");
    SIRGCN_CHECK_ARG(a->q && a->k && a->out, "q/k/out is NULL");
    SIRGCN_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->out), "q/k/out not 16-byte aligned");
    SIRGCN_CHECK_ARG(a->ldq >= ldmin && a->ldk >= ldmin && a->ldo >= ldmin, "ld smaller than the padded row (%d)", ldmin);
    SIRGCN_CHECK_ARG((a->ldq * es) % 16 == 0 && (a->ldk * es) % 16 == 0 && (a->ldo * es) % 16 == 0,
                     "row strides must be multiples of 16 bytes");
    if (mode != kFwd) {
        SIRGCN_CHECK_ARG(a->da && aligned16(a->da) && a->lda >= ldmin && (a->lda * es) % 16 == 0, "bad dA table");
    }
    if (a->e) {
        SIRGCN_CHECK_ARG(a->eid != nullptr, "edge term given without edge ids");
        SIRGCN_CHECK_ARG(aligned16(a->e) && a->lde >= ldmin && (a->lde * es) % 16 == 0, "bad e table");
    }
    if (mode == kBwdQ && a->de) {
        SIRGCN_CHECK_ARG(a->eid != nullptr, "dE requested without edge ids");
        SIRGCN_CHECK_ARG(aligned16(a->de) && a->ldde >= ldmin && (a->ldde * es) % 16 == 0, "bad dE table");
    }
    if (mode == kBwdQ && a->n_etypes > 0 && a->de_partial) {
        SIRGCN_CHECK_ARG(a->n_etypes <= SIRGCN_MAX_ETYPES && a->eid != nullptr && a->de == nullptr && aligned16(a->de_partial),
                         "table-gradient mode needs n_etypes <= %d, edge types in eid, de == NULL and aligned scratch",
                         SIRGCN_MAX_ETYPES);
    }
    if (mode == kBwdQ && a->da_scaled) {
        SIRGCN_CHECK_ARG(aligned16(a->da_scaled) && a->ldds >= ldmin && (a->ldds * es) % 16 == 0, "bad scaled-dA table");
    }
    SIRGCN_CHECK_ARG(a->n_chunks >= 0 && a->n_long >= 0 && ((a->n_chunks == 0) == (a->n_long == 0)),
                     "inconsistent schedule counts");
    SIRGCN_CHECK_ARG(a->n_tiles >= 0 && (a->n_tiles == 0 || a->tile_row), "work tiles missing (sirgcn_tiles_build)");
    if (a->n_chunks > 0) {
        SIRGCN_CHECK_ARG(a->partial && aligned16(a->partial), "partial scratch missing / misaligned");
        SIRGCN_CHECK_ARG(a->sched.long_rows && a->sched.long_first && a->sched.long_nchunks &&
                             a->sched.chunk_lrow && a->sched.chunk_beg && a->sched.big_lrows, "schedule arrays missing");
    }
    return SIRGCN_OK;
}


int dispatch(const sirgcn_edge_args *a, int mode, void *stream) {
    int rc = validate(a, mode);
    if (rc != SIRGCN_OK || a->n_rows == 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define SIRGCN_MODE(T)                                             \
    switch (mode) {                                                \
        case kFwd: return edge_launch<T, kFwd>(*a, st);            \
        case kBwdQ: return edge_launch<T, kBwdQ>(*a, st);          \
        default: return edge_launch<T, kBwdK>(*a, st);             \
    }
    switch (a->dtype) {
        case SIRGCN_F32: SIRGCN_MODE(float)
        case SIRGCN_BF16: SIRGCN_MODE(__nv_bfloat16)
        default: SIRGCN_MODE(__half)
    }
#undef SIRGCN_MODE
}

}  // namespace
}  // namespace sirgcn

extern "C" {

size_t sirgcn_edge_partial_bytes(int32_t n_chunks, int32_t d, int32_t dtype) {
    const int es = sirgcn::elem_size(dtype);
    const size_t nvec = ((size_t)d * es + 15) / 16;
    return (size_t)n_chunks * nvec * (16 / es) * sizeof(float);
}

namespace sirgcn {
namespace {
// one thread per (type, column): units summed in index order by 8 interleaved partial sums, combined in order
__global__ void __launch_bounds__(256) etable_grad_kernel(const float *__restrict__ part, int64_t n_units, int n_types,
                                                          int ldp, int d, float *__restrict__ out, int64_t ld_out) {
    __shared__ float sm[8][32];
    const int col = blockIdx.x * 32 + threadIdx.x, t = blockIdx.y, y = threadIdx.y;
    float s = 0.f;
    if (col < d)
        for (int64_t u = y; u < n_units; u += 8) s += part[(u * n_types + t) * ldp + col];
    sm[y][threadIdx.x] = s;
    __syncthreads();
    if (y == 0 && col < d) {
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) r += sm[j][threadIdx.x];
        out[(int64_t)t * ld_out + col] = r;
    }
}
}  // namespace
}  // namespace sirgcn

int sirgcn_etable_grad(const float *de_partial, int64_t n_units, int32_t n_etypes, int32_t ldp, int32_t d,
                       float *out, int64_t ld_out, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(n_units >= 0 && n_etypes > 0 && n_etypes <= SIRGCN_MAX_ETYPES && d > 0 && ldp >= d && ld_out >= d,
                     "bad n_units/n_etypes/d/ldp/ld_out");
    SIRGCN_CHECK_ARG(out && (de_partial || n_units == 0), "out/de_partial is NULL");
    etable_grad_kernel<<<dim3((unsigned)((d + 31) / 32), (unsigned)n_etypes), dim3(32, 8), 0,
                         reinterpret_cast<cudaStream_t>(stream)>>>(de_partial, n_units, n_etypes, ldp, d, out, ld_out);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_edge_fwd(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kFwd, stream); }
int sirgcn_edge_bwd_q(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kBwdQ, stream); }
int sirgcn_edge_bwd_k(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kBwdK, stream); }

}  // extern "C"
