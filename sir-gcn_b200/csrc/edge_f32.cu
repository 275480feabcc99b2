// Instantiation of the fused edge kernels for float tables (see edge_kernels.cuh).
#include "edge_kernels.cuh"
namespace sirgcn {
template int edge_launch<float>(const sirgcn_edge_args &, int, cudaStream_t);
}
