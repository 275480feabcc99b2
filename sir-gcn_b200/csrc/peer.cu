// Peer-memory transport for the row-partitioned graph (no reference counterpart — the reference
// is single-process).  One process per GPU; every rank keeps its slice of a row table (K, Q or the scaled dA) in
// a buffer that its peers map through CUDA IPC, and the all-gather of a table is (world-1) copy-engine pulls over
// NVLink/NVSwitch: no SM is taken from the edge walk that runs at the same time, and no staging copies.
// Cross-rank ordering ("every slice is written", "every pull has finished") is a flag barrier through the same
// peer mappings: one small kernel per rank, release/acquire at system scope, bounded spin.
#include "common.cuh"

#include <cstring>

namespace sirgcn {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// pads[r] = rank r's flag pad (uint32 [kPadSlots]); slot `rank` of every peer's pad <- epoch, then wait until
// every slot of the own pad has reached the epoch.  Epochs only grow (wrap-safe signed compare).
// If the spin gives up (a peer never arrived) status[0] is set to 1 and the kernel TRAPS: the context is dead, every
// later CUDA call of this process fails loudly, and no kernel ever runs on tables that were not gathered or pushed.
__global__ void peer_barrier_kernel(uint32_t *const *pads, int world, int rank, uint32_t epoch,
                                    unsigned long long timeout_ns, int *status) {
    const int t = threadIdx.x;
    __threadfence_system();          // everything this rank's earlier kernels wrote is visible to the peers
    if (t < world && t != rank) st_release_sys(pads[t] + rank, epoch);
    if (t < world && t != rank) {
        const uint32_t *mine = pads[rank] + t;
        const uint64_t t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (globaltimer_ns() - t0 > timeout_ns) {
                atomicExch(status, 1);
                __threadfence_system();
                __trap();
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

struct PushTargets {
    void *dst[SIRGCN_PEER_MAX_WORLD];
};

// out-of-place fan-out copy: every 16-byte vector of src is read once and written to each of the n_dst targets
// (peer mappings => posted writes over NVLink; the own table => a local write)
template <int UNROLL>
__global__ void __launch_bounds__(512) peer_push_kernel(const uint4 *__restrict__ src, PushTargets tg, int n_dst,
                                                         size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n_vec; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(src + i + u * stride);
        for (int t = 0; t < n_dst; ++t) {
            uint4 *d = reinterpret_cast<uint4 *>(tg.dst[t]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) stg_vec(d + i + u * stride, v[u]);
        }
    }
    for (; i < n_vec; i += stride) {
        const uint4 v = ldg_stream(src + i);
        for (int t = 0; t < n_dst; ++t) stg_vec(reinterpret_cast<uint4 *>(tg.dst[t]) + i, v);
    }
}

// ---- TMA fan-out: one elected thread per CTA moves the slice with bulk copies ----------------------------------
// global (own slice) --cp.async.bulk--> shared ring --cp.async.bulk--> every target.  No issue slots and no registers
// are spent on the data; the CTA costs its shared-memory ring (kTmaStages x kTmaChunk) and one resident warp.
constexpr int kTmaChunk = 32 * 1024;
constexpr int kTmaStages = 4;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(32) peer_push_tma_kernel(const char *__restrict__ src, PushTargets tg, int n_dst,
                                                            size_t bytes, int *status) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long full[kTmaStages];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < kTmaStages; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t n_chunks = (bytes + kTmaChunk - 1) / kTmaChunk;
    const size_t first = blockIdx.x, step = gridDim.x;
    auto chunk_bytes = [&](size_t c) { return (uint32_t)min((size_t)kTmaChunk, bytes - c * kTmaChunk); };
    auto load = [&](size_t c, int s) {
        const uint32_t bar = smem_addr(&full[s]), dst = smem_addr(ring + (size_t)s * kTmaChunk), nb = chunk_bytes(c);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src + c * kTmaChunk), "r"(nb), "r"(bar) : "memory");
    };
    // Software pipeline.  Loads run kTmaAhead chunks ahead of the stores.  The stage that chunk it+kTmaAhead is loaded
    // into was last read by the store group of chunk it+kTmaAhead-kTmaStages; when iteration `it` starts, groups up to
    // it-1 are committed, so at most kTmaStages-kTmaAhead-1 groups may still be reading their stage.
    constexpr int kTmaAhead = 2;
    static_assert(kTmaStages - kTmaAhead - 1 >= 1, "stores must be allowed to overlap");
    size_t issued = first;
    int n_issued = 0;
    for (; n_issued < kTmaAhead && issued < n_chunks; ++n_issued, issued += step) load(issued, n_issued);
    int it = 0;
    for (size_t c = first; c < n_chunks; c += step, ++it) {
        const int s = it % kTmaStages;
        const uint32_t parity = (uint32_t)(it / kTmaStages) & 1u;
        if (issued < n_chunks) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kTmaStages - kTmaAhead - 1) : "memory");
            load(issued, n_issued % kTmaStages);
            ++n_issued;
            issued += step;
        }
        unsigned spins = 0;
        while (!mbar_try_wait(smem_addr(&full[s]), parity)) {
            if (++spins > (1u << 26)) {                     // a copy that never lands: report and kill the context
                atomicExch(status, 2);
                __threadfence_system();
                __trap();
            }
        }
        const uint32_t from = smem_addr(ring + (size_t)s * kTmaChunk), nb = chunk_bytes(c);
        for (int t = 0; t < n_dst; ++t)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(reinterpret_cast<char *>(tg.dst[t]) + c * kTmaChunk), "r"(from), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // every store has been written, not only read
}

}  // namespace
}  // namespace sirgcn

extern "C" {

int sirgcn_peer_push_tma(const void *src, void *const *dsts, int32_t n_dst, size_t bytes, int32_t n_ctas,
                         int32_t *status, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(n_dst >= 1 && n_dst <= SIRGCN_PEER_MAX_WORLD && dsts && status, "bad n_dst=%d/dsts/status", n_dst);
    SIRGCN_CHECK_ARG(bytes % 16 == 0 && aligned16(src), "src/bytes must be 16-byte aligned");
    if (bytes == 0) return SIRGCN_OK;
    PushTargets tg{};
    for (int t = 0; t < n_dst; ++t) {
        SIRGCN_CHECK_ARG(dsts[t] && aligned16(dsts[t]), "target %d is NULL or unaligned", t);
        tg.dst[t] = dsts[t];
    }
    const size_t n_chunks = (bytes + kTmaChunk - 1) / kTmaChunk;
    const int grid = (int)std::min<size_t>((size_t)std::max(1, n_ctas), n_chunks);
    const int smem = kTmaStages * kTmaChunk;
    static std::atomic<bool> configured{false};
    if (!configured.load(std::memory_order_relaxed)) {
        SIRGCN_CUDA(cudaFuncSetAttribute(peer_push_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.store(true, std::memory_order_relaxed);
    }
    peer_push_tma_kernel<<<grid, 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const char *>(src), tg, n_dst, bytes, status);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_peer_push(const void *src, void *const *dsts, int32_t n_dst, size_t bytes, int32_t n_ctas, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(n_dst >= 1 && n_dst <= SIRGCN_PEER_MAX_WORLD && dsts, "bad n_dst=%d/dsts", n_dst);
    SIRGCN_CHECK_ARG(bytes % 16 == 0 && aligned16(src), "src/bytes must be 16-byte aligned");
    if (bytes == 0) return SIRGCN_OK;
    PushTargets tg{};
    for (int t = 0; t < n_dst; ++t) {
        SIRGCN_CHECK_ARG(dsts[t] && aligned16(dsts[t]), "target %d is NULL or unaligned", t);
        tg.dst[t] = dsts[t];
    }
    const size_t n_vec = bytes / 16;
    const int grid = (int)std::min<size_t>((size_t)std::max(1, n_ctas), (n_vec + 511) / 512);
    peer_push_kernel<4><<<grid, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint4 *>(src), tg, n_dst, n_vec);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

int sirgcn_peer_alloc(size_t bytes, void **dev_ptr, void *ipc_handle) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(bytes > 0 && dev_ptr && ipc_handle, "bad bytes/dev_ptr/ipc_handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == SIRGCN_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    SIRGCN_CUDA(cudaMalloc(&p, bytes));
    cudaError_t err = cudaMemset(p, 0, bytes);
    if (err == cudaSuccess) err = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(ipc_handle), p);
    if (err != cudaSuccess) {
        cudaFree(p);
        set_error("peer_alloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
        return (int)err;
    }
    *dev_ptr = p;
    return SIRGCN_OK;
}

int sirgcn_peer_free(void *dev_ptr) {
    using namespace sirgcn;
    if (dev_ptr) SIRGCN_CUDA(cudaFree(dev_ptr));
    return SIRGCN_OK;
}

int sirgcn_peer_open(const void *ipc_handle, void **dev_ptr) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(ipc_handle && dev_ptr, "bad ipc_handle/dev_ptr");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    SIRGCN_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SIRGCN_OK;
}

int sirgcn_peer_close(void *dev_ptr) {
    using namespace sirgcn;
    if (dev_ptr) SIRGCN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return SIRGCN_OK;
}

int sirgcn_peer_copy(void *dst, const void *src, size_t bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(bytes == 0 || (dst && src), "bad dst/src");
    if (bytes == 0) return SIRGCN_OK;
    // device-to-device between two mappings of different GPUs: the driver routes it to a copy engine over NVLink
    SIRGCN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return SIRGCN_OK;
}

int sirgcn_peer_barrier(uint32_t *const *pads, int32_t world, int32_t rank, uint32_t epoch, uint64_t timeout_ns,
                        int32_t *status, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(pads && status && world >= 1 && world <= SIRGCN_PEER_MAX_WORLD && rank >= 0 && rank < world,
                     "bad pads/status/world=%d/rank=%d", world, rank);
    if (world == 1) return SIRGCN_OK;
    peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pads, world, rank, epoch,
                                                                             (unsigned long long)timeout_ns, status);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

}  // extern "C"
