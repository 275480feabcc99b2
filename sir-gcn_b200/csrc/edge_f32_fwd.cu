// Instantiation of the fused edge kernels: float tables, kFwd walk (see edge_kernels.cuh).
#include "edge_kernels.cuh"
namespace sirgcn {
template int edge_launch<float, kFwd>(const sirgcn_edge_args &, cudaStream_t);
}
