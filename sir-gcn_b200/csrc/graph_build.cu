// Graph index construction on the GPU: COO -> destination-major CSR + source-major CSC (stable),
// degree norms and the long-row schedule.  Replaces DGL's format conversion and degree queries
// (/root/reference/models/conv.py:51-52 in_degrees/out_degrees, :63 update_all over the in-CSR).
// The stable LSD radix sort is cub::DeviceRadixSort (ships with the CUDA toolkit); everything
// else is hand-written.  Bit-exact against oracle/csr_ref.c and torch.sort(stable=True).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace sirgcn {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(int64_t n, int per_block = kThreads) {
    return (unsigned)((n + per_block - 1) / per_block);
}

// keys outside [0, num_rows) are clamped (nothing downstream may index out of bounds) and reported through *bad:
// the host raises after its one sync per graph (DGL raises for such a graph too)
__global__ void iota_copy_kernel(const int32_t *__restrict__ key_in, int32_t *__restrict__ key_out,
                                 int32_t *__restrict__ val_out, int64_t n, int32_t num_rows, int32_t *bad) {
    bool any_bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t k = key_in[i];
        if (k < 0 || k >= num_rows) {
            any_bad = true;
            k = k < 0 ? 0 : num_rows - 1;
        }
        key_out[i] = k;
        val_out[i] = (int32_t)i;
    }
    if (any_bad && bad) atomicOr(bad, 1);
}

// indptr[v] = first position whose key is >= v, from the sorted key array (run boundaries).
__global__ void indptr_from_sorted_kernel(const int32_t *__restrict__ keys, int64_t n, int32_t num_nodes,
                                          int32_t *__restrict__ indptr) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t prev = i == 0 ? -1 : keys[i - 1];
        const int32_t cur = i == n ? num_nodes : keys[i];
        for (int32_t v = prev + 1; v <= cur; ++v) indptr[v] = (int32_t)i;
    }
}

__global__ void empty_indptr_kernel(int32_t *indptr, int32_t num_nodes) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= num_nodes; i += gridDim.x * blockDim.x) indptr[i] = 0;
}

__global__ void gather_i32_kernel(const int32_t *__restrict__ table, const int32_t *__restrict__ sel,
                                  int32_t *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = table[sel[i]];
}

__global__ void copy_i32_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

__global__ void norms_kernel(const int32_t *__restrict__ indptr_in, const int32_t *__restrict__ indptr_out,
                             int32_t num_nodes, float *in_norm, float *out_norm, float *inv_in_deg) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < num_nodes; v += gridDim.x * blockDim.x) {
        const float di = (float)max(indptr_in[v + 1] - indptr_in[v], 1);
        const float dq = (float)max(indptr_out[v + 1] - indptr_out[v], 1);
        // IEEE sqrt and division: identical to the CPU oracle's 1.0f / sqrtf(x)
        if (in_norm) in_norm[v] = __fdiv_rn(1.f, __fsqrt_rn(di));
        if (out_norm) out_norm[v] = __fdiv_rn(1.f, __fsqrt_rn(dq));
        if (inv_in_deg) inv_in_deg[v] = __fdiv_rn(1.f, di);
    }
}

// Rows with degree > thr are "long": split into ceil(deg/thr) chunks.  Slots are claimed with
// integer atomics; slot numbering does not influence any floating-point result because each
// row's partials are always summed in chunk order.
__global__ void schedule_kernel(const int32_t *__restrict__ indptr, int32_t num_rows, int32_t thr,
                                sirgcn_schedule s, int32_t *counts /* {n_long, n_chunks} */) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < num_rows; row += gridDim.x * blockDim.x) {
        const int beg = indptr[row];
        const int deg = indptr[row + 1] - beg;
        if (deg <= thr) continue;
        const int nch = (deg + thr - 1) / thr;
        const int l = atomicAdd(&counts[0], 1);
        const int first = atomicAdd(&counts[1], nch);
        s.long_rows[l] = row;
        s.long_first[l] = first;
        s.long_nchunks[l] = nch;
        if (nch > SIRGCN_BIG_CHUNKS) s.big_lrows[1 + atomicAdd(&s.big_lrows[0], 1)] = l;
        for (int ch = 0; ch < nch; ++ch) {
            s.chunk_lrow[first + ch] = l;
            s.chunk_beg[first + ch] = beg + ch * thr;
        }
    }
}

// tile_row[t] = min { r : indptr[r] + ROW_COST*r >= t*TILE_WORK }  (see sirgcn.h "Work tiles")
__global__ void tiles_kernel(const int32_t *__restrict__ indptr, int32_t num_rows, int64_t n_tiles,
                             int32_t *__restrict__ tile_row) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= num_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = (int64_t)indptr[r] + (int64_t)SIRGCN_ROW_COST * r;
        const int64_t t_hi = w / SIRGCN_TILE_WORK;
        const int64_t t_lo = r == 0 ? 0 : ((int64_t)indptr[r - 1] + (int64_t)SIRGCN_ROW_COST * (r - 1)) / SIRGCN_TILE_WORK + 1;
        for (int64_t t = t_lo; t <= t_hi && t < n_tiles; ++t) tile_row[t] = (int32_t)r;
        if (r == num_rows) tile_row[n_tiles] = num_rows;
    }
}

// ---- edge-subset view of a converted graph (no sort: compaction keeps the parent's stable order) ---------------
__global__ void keep_flags_kernel(const uint8_t *__restrict__ keep, const int32_t *__restrict__ eid,
                                  int32_t *__restrict__ flag, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = keep[eid ? eid[i] : i] ? 1 : 0;
}

// positions: prefix[p] = number of kept positions before p (exclusive scan of the flags, prefix[n] = total)
__global__ void compact_rows_kernel(const uint8_t *__restrict__ keep, const int32_t *__restrict__ eid,
                                    const int32_t *__restrict__ other, const int32_t *__restrict__ prefix,
                                    const int32_t *__restrict__ new_id, int32_t *__restrict__ other_out,
                                    int32_t *__restrict__ eid_out, int64_t n) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = eid[p];
        if (!keep[e]) continue;
        const int32_t q = prefix[p];
        other_out[q] = other[p];
        if (eid_out) eid_out[q] = new_id[e];
    }
}

__global__ void compact_indptr_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ prefix,
                                      int32_t num_rows, int32_t *__restrict__ indptr_out) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= num_rows; r += gridDim.x * blockDim.x)
        indptr_out[r] = prefix[indptr[r]];
}

struct SortScratch {
    int32_t *keys[2];
    int32_t *vals[2];
    void *cub_tmp;
    size_t cub_bytes;
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

size_t cub_temp_bytes(int64_t E, int end_bit) {
    size_t bytes = 0;
    cub::DoubleBuffer<int32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, E, 0, end_bit);
    return bytes;
}

int key_bits(int32_t num_nodes) {
    int b = 1;
    while (b < 31 && (1LL << b) < (long long)num_nodes) ++b;
    return b;
}

// sort (key, edge id) pairs; write indptr, the "other endpoint" array and (optionally) edge ids
int build_one(const int32_t *key, const int32_t *other, int64_t E, int32_t N, int32_t *indptr,
              int32_t *other_sorted, int32_t *eid_sorted, SortScratch &ws, cudaStream_t st, int32_t *bad = nullptr) {
    if (E == 0) {
        empty_indptr_kernel<<<blocks_for(N + 1), kThreads, 0, st>>>(indptr, N);
        SIRGCN_LAUNCHED();
        return SIRGCN_OK;
    }
    const unsigned grid = std::min(blocks_for(E), (unsigned)kNumSMs * 16);
    iota_copy_kernel<<<grid, kThreads, 0, st>>>(key, ws.keys[0], ws.vals[0], E, N, bad);
    SIRGCN_LAUNCHED();
    cub::DoubleBuffer<int32_t> k(ws.keys[0], ws.keys[1]), v(ws.vals[0], ws.vals[1]);
    size_t bytes = ws.cub_bytes;
    SIRGCN_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, bytes, k, v, E, 0, key_bits(N), st));
    g_launches.fetch_add(1);
    indptr_from_sorted_kernel<<<grid, kThreads, 0, st>>>(k.Current(), E, N, indptr);
    SIRGCN_LAUNCHED();
    gather_i32_kernel<<<grid, kThreads, 0, st>>>(other, v.Current(), other_sorted, E);
    SIRGCN_LAUNCHED();
    if (eid_sorted) {
        copy_i32_kernel<<<grid, kThreads, 0, st>>>(v.Current(), eid_sorted, E);
        SIRGCN_LAUNCHED();
    }
    return SIRGCN_OK;
}

int check_sched(const sirgcn_schedule *s) {
    SIRGCN_CHECK_ARG(s && s->long_rows && s->long_first && s->long_nchunks && s->chunk_lrow && s->chunk_beg && s->big_lrows,
                     "schedule arrays missing");
    return SIRGCN_OK;
}

}  // namespace
}  // namespace sirgcn

extern "C" {

const char *sirgcn_last_error(void) { return sirgcn::g_err; }
int sirgcn_abi_version(void) { return SIRGCN_ABI_VERSION; }
uint64_t sirgcn_launch_count(void) { return sirgcn::g_launches.load(); }

size_t sirgcn_csr_build_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
    using namespace sirgcn;
    if (num_edges <= 0) return 256;
    const size_t arr = align_up((size_t)num_edges * sizeof(int32_t));
    return 4 * arr + align_up(cub_temp_bytes(num_edges, key_bits(num_nodes))) + 256;
}

int64_t sirgcn_num_tiles(int32_t num_rows, int64_t num_edges) {
    if (num_rows <= 0) return 0;
    return (num_edges + (int64_t)SIRGCN_ROW_COST * num_rows) / SIRGCN_TILE_WORK + 1;
}

int sirgcn_tiles_build(const int32_t *indptr, int32_t num_rows, int64_t num_edges, int32_t *tile_row, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_rows >= 0 && num_edges >= 0, "bad num_rows/num_edges");
    if (num_rows == 0) return SIRGCN_OK;
    SIRGCN_CHECK_ARG(indptr && tile_row, "indptr/tile_row is NULL");
    const int64_t n_tiles = sirgcn_num_tiles(num_rows, num_edges);
    SIRGCN_CHECK_ARG(n_tiles < (1LL << 31), "too many tiles");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    tiles_kernel<<<std::min(blocks_for((int64_t)num_rows + 1), (unsigned)kNumSMs * 16), kThreads, 0, st>>>(
        indptr, num_rows, n_tiles, tile_row);
    SIRGCN_LAUNCHED();
    return SIRGCN_OK;
}

size_t sirgcn_edge_subgraph_workspace_bytes(int64_t num_edges) {
    using namespace sirgcn;
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (int32_t *)nullptr, (int32_t *)nullptr, num_edges + 1);
    // flags [E+1] + prefix [E+1] + cub scratch
    return 2 * align_up((size_t)(num_edges + 1) * sizeof(int32_t)) + align_up(scan) + 256;
}

int sirgcn_edge_subgraph(const uint8_t *keep, int64_t num_edges, int32_t num_nodes,
                         const int32_t *indptr_in, const int32_t *col_src, const int32_t *eid_in,
                         const int32_t *indptr_out, const int32_t *row_dst, const int32_t *eid_out,
                         int32_t *new_id, int32_t *sub_indptr_in, int32_t *sub_col_src, int32_t *sub_eid_in,
                         int32_t *sub_indptr_out, int32_t *sub_row_dst, int32_t *sub_eid_out,
                         float *in_norm, float *out_norm, float *inv_in_deg,
                         void *workspace, size_t workspace_bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_edges >= 0 && num_edges < (1LL << 31) && num_nodes >= 0, "bad num_edges/num_nodes");
    SIRGCN_CHECK_ARG(indptr_in && indptr_out && sub_indptr_in && sub_indptr_out && new_id, "indptr/new_id arrays missing");
    SIRGCN_CHECK_ARG(num_edges == 0 || (keep && col_src && eid_in && row_dst && eid_out && sub_col_src && sub_row_dst),
                     "edge arrays missing (the parent graph must carry edge ids)");
    SIRGCN_CHECK_ARG(workspace && workspace_bytes >= sirgcn_edge_subgraph_workspace_bytes(num_edges), "workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t E = num_edges;
    char *base = reinterpret_cast<char *>(workspace);
    int32_t *flag = reinterpret_cast<int32_t *>(base);
    int32_t *prefix = reinterpret_cast<int32_t *>(base + align_up((size_t)(E + 1) * sizeof(int32_t)));
    void *cub_tmp = base + 2 * align_up((size_t)(E + 1) * sizeof(int32_t));
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, flag, prefix, E + 1);
    const unsigned grid = std::max(1u, std::min(blocks_for(E + 1), (unsigned)kNumSMs * 16));
    const unsigned grid_n = std::max(1u, std::min(blocks_for(num_nodes + 1), (unsigned)kNumSMs * 16));
    SIRGCN_CUDA(cudaMemsetAsync(flag, 0, (size_t)(E + 1) * sizeof(int32_t), st));
    // 1. new edge ids: rank of every kept edge among the kept ones (edge-id order is preserved)
    if (E) {
        keep_flags_kernel<<<grid, kThreads, 0, st>>>(keep, nullptr, flag, E);
        SIRGCN_LAUNCHED();
    }
    SIRGCN_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, scan, flag, prefix, E + 1, st));
    g_launches.fetch_add(1);
    SIRGCN_CUDA(cudaMemcpyAsync(new_id, prefix, (size_t)(E + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    // 2./3. both compressed-row structures: flags in stored order -> exclusive scan -> compaction
    const int32_t *ips[2] = {indptr_in, indptr_out}, *others[2] = {col_src, row_dst}, *eids[2] = {eid_in, eid_out};
    int32_t *sips[2] = {sub_indptr_in, sub_indptr_out}, *sothers[2] = {sub_col_src, sub_row_dst},
            *seids[2] = {sub_eid_in, sub_eid_out};
    for (int s = 0; s < 2; ++s) {
        if (E) {
            keep_flags_kernel<<<grid, kThreads, 0, st>>>(keep, eids[s], flag, E);
            SIRGCN_LAUNCHED();
        }
        SIRGCN_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, scan, flag, prefix, E + 1, st));
        g_launches.fetch_add(1);
        if (E) {
            compact_rows_kernel<<<grid, kThreads, 0, st>>>(keep, eids[s], others[s], prefix, new_id, sothers[s], seids[s], E);
            SIRGCN_LAUNCHED();
        }
        compact_indptr_kernel<<<grid_n, kThreads, 0, st>>>(ips[s], prefix, num_nodes, sips[s]);
        SIRGCN_LAUNCHED();
    }
    if ((in_norm || out_norm || inv_in_deg) && num_nodes > 0) {
        norms_kernel<<<grid_n, kThreads, 0, st>>>(sub_indptr_in, sub_indptr_out, num_nodes, in_norm, out_norm, inv_in_deg);
        SIRGCN_LAUNCHED();
    }
    return SIRGCN_OK;
}

int sirgcn_schedule_build(const int32_t *indptr, int32_t num_rows, int32_t long_threshold,
                          const sirgcn_schedule *sched, int32_t *counts, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(indptr && counts, "indptr/counts is NULL");
    SIRGCN_CHECK_ARG(long_threshold >= 32, "long_threshold must be >= 32");
    int rc = check_sched(sched);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SIRGCN_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
    SIRGCN_CUDA(cudaMemsetAsync(sched->big_lrows, 0, sizeof(int32_t), st));
    if (num_rows > 0) {
        schedule_kernel<<<std::min(blocks_for(num_rows), (unsigned)kNumSMs * 16), kThreads, 0, st>>>(
            indptr, num_rows, long_threshold, *sched, counts);
        SIRGCN_LAUNCHED();
    }
    return SIRGCN_OK;
}

size_t sirgcn_rows_build_workspace_bytes(int64_t num_pos, int32_t num_rows) {
    return sirgcn_csr_build_workspace_bytes(num_pos, num_rows);
}

int sirgcn_rows_build(const int32_t *key, const int32_t *other, int64_t num_pos, int32_t num_rows,
                      int32_t *indptr, int32_t *other_sorted, int32_t *eid_sorted,
                      void *workspace, size_t workspace_bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_pos >= 0 && num_pos < (1LL << 31), "num_pos=%lld out of int32 range", (long long)num_pos);
    SIRGCN_CHECK_ARG(num_rows >= 0 && indptr, "bad num_rows / indptr");
    SIRGCN_CHECK_ARG(num_pos == 0 || (key && other && other_sorted), "key/other arrays are NULL");
    const size_t need = sirgcn_rows_build_workspace_bytes(num_pos, num_rows);
    if (workspace_bytes < need || (!workspace && num_pos > 0)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return SIRGCN_ENOSPC;
    }
    SortScratch ws{};
    if (num_pos > 0) {
        const size_t arr = align_up((size_t)num_pos * sizeof(int32_t));
        char *base = reinterpret_cast<char *>(workspace);
        ws.keys[0] = reinterpret_cast<int32_t *>(base);
        ws.keys[1] = reinterpret_cast<int32_t *>(base + arr);
        ws.vals[0] = reinterpret_cast<int32_t *>(base + 2 * arr);
        ws.vals[1] = reinterpret_cast<int32_t *>(base + 3 * arr);
        ws.cub_tmp = base + 4 * arr;
        ws.cub_bytes = workspace_bytes - 4 * arr;
    }
    return build_one(key, other, num_pos, num_rows, indptr, other_sorted, eid_sorted, ws, reinterpret_cast<cudaStream_t>(stream));
}

int sirgcn_csr_build(const int32_t *src, const int32_t *dst, int64_t num_edges, int32_t num_nodes,
                     int32_t *indptr_in, int32_t *col_src, int32_t *eid_in,
                     int32_t *indptr_out, int32_t *row_dst, int32_t *eid_out,
                     float *in_norm, float *out_norm, float *inv_in_deg,
                     int32_t long_threshold, const sirgcn_schedule *sched_in, const sirgcn_schedule *sched_out,
                     int32_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace sirgcn;
    SIRGCN_CHECK_ARG(num_edges >= 0 && num_edges < (1LL << 31), "num_edges=%lld out of int32 range", (long long)num_edges);
    SIRGCN_CHECK_ARG(num_nodes >= 0, "num_nodes < 0");
    SIRGCN_CHECK_ARG(indptr_in && indptr_out, "indptr outputs are NULL");
    SIRGCN_CHECK_ARG(num_edges == 0 || (src && dst && col_src && row_dst), "edge arrays are NULL");
    const size_t need = sirgcn_csr_build_workspace_bytes(num_edges, num_nodes);
    if (workspace_bytes < need || (!workspace && num_edges > 0)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return SIRGCN_ENOSPC;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    SortScratch ws{};
    if (num_edges > 0) {
        const size_t arr = align_up((size_t)num_edges * sizeof(int32_t));
        char *base = reinterpret_cast<char *>(workspace);
        ws.keys[0] = reinterpret_cast<int32_t *>(base);
        ws.keys[1] = reinterpret_cast<int32_t *>(base + arr);
        ws.vals[0] = reinterpret_cast<int32_t *>(base + 2 * arr);
        ws.vals[1] = reinterpret_cast<int32_t *>(base + 3 * arr);
        ws.cub_tmp = base + 4 * arr;
        ws.cub_bytes = workspace_bytes - 4 * arr;
    }
    SIRGCN_CHECK_ARG(num_edges == 0 || num_nodes > 0, "edges on a graph without nodes");
    int32_t *bad = counts ? counts + 4 : nullptr;        // every id is a key of one of the two sorts: both are checked
    if (bad) SIRGCN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    int rc = build_one(dst, src, num_edges, num_nodes, indptr_in, col_src, eid_in, ws, st, bad);
    if (rc) return rc;
    rc = build_one(src, dst, num_edges, num_nodes, indptr_out, row_dst, eid_out, ws, st, bad);
    if (rc) return rc;
    if ((in_norm || out_norm || inv_in_deg) && num_nodes > 0) {
        norms_kernel<<<std::min(blocks_for(num_nodes), (unsigned)kNumSMs * 16), kThreads, 0, st>>>(
            indptr_in, indptr_out, num_nodes, in_norm, out_norm, inv_in_deg);
        SIRGCN_LAUNCHED();
    }
    if (counts) {
        rc = sirgcn_schedule_build(indptr_in, num_nodes, long_threshold, sched_in, counts, stream);
        if (rc) return rc;
        rc = sirgcn_schedule_build(indptr_out, num_nodes, long_threshold, sched_out, counts + 2, stream);
        if (rc) return rc;
    }
    return SIRGCN_OK;
}

}  // extern "C"
