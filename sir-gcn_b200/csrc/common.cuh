// Shared device/host helpers for libsirgcn (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <algorithm>

#include "../../include/sirgcn.h"

namespace sirgcn {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define SIRGCN_CHECK_ARG(cond, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            sirgcn::set_error(__VA_ARGS__);  \
            return SIRGCN_EINVAL;            \
        }                                    \
    } while (0)

#define SIRGCN_CUDA(call)                                                                   \
    do {                                                                                    \
        cudaError_t err__ = (call);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            sirgcn::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),    \
                              __FILE__, __LINE__);                                          \
            return (int)err__;                                                              \
        }                                                                                   \
    } while (0)

// call after every <<<>>> launch
#define SIRGCN_LAUNCHED()                    \
    do {                                     \
        sirgcn::g_launches.fetch_add(1);     \
        SIRGCN_CUDA(cudaGetLastError());     \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int elem_size(int dtype) { return dtype == SIRGCN_F32 ? 4 : 2; }

constexpr int kNumSMs = 148;  // B200

// ---- 16-byte vectors of table elements ---------------------------------------------------
template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int N = 4; };
template <> struct VecTraits<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct VecTraits<__half> { static constexpr int N = 8; };

// read-only 128-bit gather that does not allocate in L1 (rows are touched once per warp)
__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_keep(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg_vec(void *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }

template <typename T> __device__ __forceinline__ void unpack(const uint4 &raw, float (&f)[VecTraits<T>::N]);
template <> __device__ __forceinline__ void unpack<float>(const uint4 &raw, float (&f)[4]) {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
    f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4 &raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // bf16 -> fp32 is a 16-bit shift
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void unpack<__half>(const uint4 &raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = *reinterpret_cast<const __half2 *>(&w[i]);
        float2 t = __half22float2(h);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

template <typename T> __device__ __forceinline__ uint4 pack(const float (&f)[VecTraits<T>::N]);
template <> __device__ __forceinline__ uint4 pack<float>(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack<__half>(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- activations (fp32, precise: parity target is 1e-5 relative) --------------------------
__device__ __forceinline__ float act_fwd(float z, int act, float p) {
    switch (act) {
        case SIRGCN_ACT_RELU: return z > 0.f ? z : 0.f;
        case SIRGCN_ACT_LEAKY_RELU: return z > 0.f ? z : p * z;
        case SIRGCN_ACT_GELU: return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
        default: return z;
    }
}
// derivative wrt the pre-activation (matches ATen: relu' (0) = 0, leaky_relu'(0) = slope)
__device__ __forceinline__ float act_bwd(float z, int act, float p) {
    switch (act) {
        case SIRGCN_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case SIRGCN_ACT_LEAKY_RELU: return z > 0.f ? 1.f : p;
        case SIRGCN_ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
            return cdf + z * pdf;
        }
        default: return 1.f;
    }
}

}  // namespace sirgcn
