// Instantiation of the fused edge kernels for __nv_bfloat16 tables (see edge_kernels.cuh).
#include "edge_kernels.cuh"
namespace sirgcn {
template int edge_launch<__nv_bfloat16>(const sirgcn_edge_args &, int, cudaStream_t);
}
