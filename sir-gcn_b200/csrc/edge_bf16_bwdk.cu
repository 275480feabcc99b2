// Instantiation of the fused edge kernels: __nv_bfloat16 tables, kBwdK walk (see edge_kernels.cuh).
#include "edge_kernels.cuh"
namespace sirgcn {
template int edge_launch<__nv_bfloat16, kBwdK>(const sirgcn_edge_args &, cudaStream_t);
}
