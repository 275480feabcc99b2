// Instantiation of the fused edge kernels: __nv_bfloat16 tables, kFwd walk (see edge_kernels.cuh).
#include "edge_kernels.cuh"
namespace sirgcn {
template int edge_launch<__nv_bfloat16, kFwd>(const sirgcn_edge_args &, cudaStream_t);
}
