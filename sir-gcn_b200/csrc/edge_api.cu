// C-ABI entry points of the fused edge stage (argument validation + dtype dispatch).
#include "common.cuh"

namespace sirgcn {
enum Mode { kFwd = 0, kBwdQ = 1, kBwdK = 2 };
template <typename T> int edge_launch(const sirgcn_edge_args &a, int mode, cudaStream_t st);
extern template int edge_launch<float>(const sirgcn_edge_args &, int, cudaStream_t);
extern template int edge_launch<__nv_bfloat16>(const sirgcn_edge_args &, int, cudaStream_t);
extern template int edge_launch<__half>(const sirgcn_edge_args &, int, cudaStream_t);

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
    SIRGCN_CHECK_ARG(a->n_chunks >= 0 && a->n_long >= 0 && ((a->n_chunks == 0) == (a->n_long == 0)),
                     "inconsistent schedule counts");
    if (a->n_chunks > 0) {
        SIRGCN_CHECK_ARG(a->partial && aligned16(a->partial), "partial scratch missing / misaligned");
        SIRGCN_CHECK_ARG(a->sched.long_rows && a->sched.long_first && a->sched.long_nchunks &&
                             a->sched.chunk_lrow && a->sched.chunk_beg, "schedule arrays missing");
    }
    return SIRGCN_OK;
}


int dispatch(const sirgcn_edge_args *a, int mode, void *stream) {
    int rc = validate(a, mode);
    if (rc != SIRGCN_OK || a->n_rows == 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (a->dtype) {
        case SIRGCN_F32: return edge_launch<float>(*a, mode, st);
        case SIRGCN_BF16: return edge_launch<__nv_bfloat16>(*a, mode, st);
        default: return edge_launch<__half>(*a, mode, st);
    }
}

}  // namespace
}  // namespace sirgcn

extern "C" {

size_t sirgcn_edge_partial_bytes(int32_t n_chunks, int32_t d, int32_t dtype) {
    const int es = sirgcn::elem_size(dtype);
    const size_t nvec = ((size_t)d * es + 15) / 16;
    return (size_t)n_chunks * nvec * (16 / es) * sizeof(float);
}

int sirgcn_edge_fwd(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kFwd, stream); }
int sirgcn_edge_bwd_q(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kBwdQ, stream); }
int sirgcn_edge_bwd_k(const sirgcn_edge_args *args, void *stream) { return sirgcn::dispatch(args, sirgcn::kBwdK, stream); }

}  // extern "C"
