/* sirgcn.h — C-ABI of the B200-native SIR-GCN convolution library (libsirgcn.so).
 *
 * The reference (briangodwinlim/SIR-GCN) has no FFI layer of its own: its hot path is
 * models/conv.py executed by DGL + PyTorch.  Each entry point below names the reference
 * call site (file:line under /root/reference) whose library behaviour it replaces.
 *
 * Conventions
 *  - every pointer is a raw DEVICE pointer owned by the caller (PyTorch caching
 *    allocator); the library never allocates or frees tensor memory and never
 *    synchronises the device (exception: none — counts needed on the host are written
 *    to device memory and the caller decides when to read them);
 *  - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*);
 *  - return value 0 = success, otherwise a negative SIRGCN_E* code or a positive
 *    cudaError_t; sirgcn_last_error() returns a thread-local message;
 *  - feature tables are row-major, `ld*` = row stride in ELEMENTS; every row start
 *    must be 16-byte aligned (base pointer 16-B aligned, ld*sizeof(elem) % 16 == 0);
 *    columns d..ld-1 of Q/K/E tables must hold zeros (they are read as padding);
 *  - indices are int32; E < 2^31.
 */
#ifndef SIRGCN_H_
#define SIRGCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIRGCN_ABI_VERSION 9

/* element types of feature tables (accumulation is always fp32) */
enum { SIRGCN_F32 = 0, SIRGCN_BF16 = 1, SIRGCN_F16 = 2 };
/* activations σ applied in registers (conv.py:34 `activation`) */
enum { SIRGCN_ACT_IDENTITY = 0, SIRGCN_ACT_RELU = 1, SIRGCN_ACT_LEAKY_RELU = 2, SIRGCN_ACT_GELU = 3 };
/* error codes */
enum {
    SIRGCN_OK = 0,
    SIRGCN_EINVAL = -1,   /* bad argument (shape, alignment, enum)          */
    SIRGCN_ENOSPC = -2,   /* workspace too small                            */
    SIRGCN_EUNSUP = -3    /* unsupported configuration (e.g. row too wide)  */
};

const char *sirgcn_last_error(void);
int sirgcn_abi_version(void);
/* number of kernels this library has launched in this process (bench `gpu_launches`) */
uint64_t sirgcn_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Graph index construction.  Replaces DGL's COO->CSR/CSC conversion and degree queries
 * triggered by graph.in_degrees()/out_degrees() (conv.py:51-52) and update_all (conv.py:63).
 *
 * From COO (src[e] -> dst[e], e = edge id) build, by STABLE sort (ties keep edge-id order):
 *   in-CSR  (destination-major): indptr_in[N+1], col_src[E], eid_in[E]
 *   out-CSC (source-major)     : indptr_out[N+1], row_dst[E], eid_out[E]
 * eid_in / eid_out may be NULL (then no edge-id permutation is produced).
 * Per-node coefficients (any may be NULL):
 *   in_norm[v]  = clamp(in_deg,1)^-1/2,  out_norm[v] = clamp(out_deg,1)^-1/2  (conv.py:54-57, 'sym')
 *   inv_in_deg[v] = 1/clamp(in_deg,1)                                         (fn.mean, conv.py:41)
 * Long-row schedule (rows with degree > long_threshold are split into chunks of
 * long_threshold edges so that hubs are spread over many warps):
 *   sched_*_long_rows[cap_long], sched_*_long_first[cap_long + 1 ... see below],
 *   sched_*_chunk_lrow[cap_chunks], sched_*_chunk_beg[cap_chunks], counts[5] =
 *   {n_long_in, n_chunks_in, n_long_out, n_chunks_out, bad_ids} (device int32).  bad_ids != 0: some src/dst id was
 *   outside [0, num_nodes) — such ids are clamped so that nothing indexes out of bounds, and the caller must
 *   reject the graph once it has read the counts (DGL raises for such a graph as well).
 *   cap_long = E/long_threshold + 1, cap_chunks = 2*E/long_threshold + 2.
 * workspace: sirgcn_csr_build_workspace_bytes(E, N) bytes of device scratch.
 */
size_t sirgcn_csr_build_workspace_bytes(int64_t num_edges, int32_t num_nodes);

typedef struct sirgcn_schedule {
    int32_t *long_rows;    /* [cap_long]   row id of each long row                        */
    int32_t *long_first;   /* [cap_long]   first chunk slot of each long row              */
    int32_t *long_nchunks; /* [cap_long]   number of chunks of each long row              */
    int32_t *chunk_lrow;   /* [cap_chunks] index into long_rows for each chunk            */
    int32_t *chunk_beg;    /* [cap_chunks] first edge position of each chunk              */
    int32_t *big_lrows;    /* [cap_big + 1] [0] = number of long rows with more than SIRGCN_BIG_CHUNKS chunks,
                              then their indices into long_rows (hubs get a whole CTA in the ordered
                              final sum); cap_big = E / (long_threshold * SIRGCN_BIG_CHUNKS) + 2       */
} sirgcn_schedule;
#define SIRGCN_BIG_CHUNKS 32

int sirgcn_csr_build(const int32_t *src, const int32_t *dst, int64_t num_edges, int32_t num_nodes,
                     int32_t *indptr_in, int32_t *col_src, int32_t *eid_in,
                     int32_t *indptr_out, int32_t *row_dst, int32_t *eid_out,
                     float *in_norm, float *out_norm, float *inv_in_deg,
                     int32_t long_threshold,
                     const sirgcn_schedule *sched_in, const sirgcn_schedule *sched_out,
                     int32_t *counts /* [5] device */,
                     void *workspace, size_t workspace_bytes, void *stream);

/* One compressed-row structure from (key, other) pairs: rows = key values in [0, num_rows), stable
 * (ties keep input order); `other` values are carried along untouched (they may be GLOBAL node ids of
 * a row-partitioned graph, partition.py).  eid_sorted (input position of every stored entry) may be NULL. */
size_t sirgcn_rows_build_workspace_bytes(int64_t num_pos, int32_t num_rows);
int sirgcn_rows_build(const int32_t *key, const int32_t *other, int64_t num_pos, int32_t num_rows,
                      int32_t *indptr, int32_t *other_sorted, int32_t *eid_sorted,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Edge-subset view of a converted graph, WITHOUT sorting: replaces the per-layer, per-step graph rebuild behind
 * DropEdge (models/utils.py:96-102, used by zinc/model.py:50, super-pixel/model.py:44, sbm-dataset/model.py:42,
 * wiki-cs/model.py:41), where DGL's remove_edges makes a new graph whose formats are converted again.
 * keep[e] != 0 marks the edges of the parent (by edge id) that stay.  Kept edges are renumbered in edge-id order:
 * new_id[e] = number of kept edges with id < e (new_id has num_edges + 1 entries, new_id[num_edges] = number kept;
 * the caller reads it to size the views).  Because the parent's orderings are stable, compaction of its in-CSR and
 * out-CSC gives exactly the structures sirgcn_csr_build would produce from the kept COO edges (bit-exact, tested).
 * sub_* arrays are sized for num_edges entries (the first new_id[num_edges] are written); sub_eid_* may be NULL.
 * The parent must carry edge ids (eid_in / eid_out). */
size_t sirgcn_edge_subgraph_workspace_bytes(int64_t num_edges);
int sirgcn_edge_subgraph(const uint8_t *keep, int64_t num_edges, int32_t num_nodes,
                         const int32_t *indptr_in, const int32_t *col_src, const int32_t *eid_in,
                         const int32_t *indptr_out, const int32_t *row_dst, const int32_t *eid_out,
                         int32_t *new_id /* [num_edges + 1] */,
                         int32_t *sub_indptr_in, int32_t *sub_col_src, int32_t *sub_eid_in,
                         int32_t *sub_indptr_out, int32_t *sub_row_dst, int32_t *sub_eid_out,
                         float *in_norm, float *out_norm, float *inv_in_deg,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Schedule only (graph already in CSR form, e.g. generated on device): fills one
 * sirgcn_schedule and counts[0..1] = {n_long, n_chunks}. */
int sirgcn_schedule_build(const int32_t *indptr, int32_t num_rows, int32_t long_threshold,
                          const sirgcn_schedule *sched, int32_t *counts /* [2] device */, void *stream);

/* Work tiles: rows are grouped so that every tile holds about SIRGCN_TILE_WORK work units, one
 * unit per stored edge plus SIRGCN_ROW_COST per row (the q read / A write of a row costs about as
 * much as a few gathered neighbours), whatever the degree distribution.  With
 * w(r) = indptr[r] + SIRGCN_ROW_COST * r, tile t holds the rows with floor(w(r)/SIRGCN_TILE_WORK) == t:
 *   tile_row[t] = min { r : w(r) >= t * SIRGCN_TILE_WORK },   tile_row[n_tiles] = num_rows,
 *   n_tiles = sirgcn_num_tiles(num_rows, num_edges) = (num_edges + ROW_COST*num_rows) / TILE_WORK + 1.
 * A tile never holds more than TILE_WORK / ROW_COST = 64 rows.  Purely a function of indptr, so
 * the decomposition — and with it every floating-point summation order — is reproducible. */
#ifndef SIRGCN_TILE_WORK
#define SIRGCN_TILE_WORK 256
#endif
#define SIRGCN_ROW_COST 4
int64_t sirgcn_num_tiles(int32_t num_rows, int64_t num_edges);
int sirgcn_tiles_build(const int32_t *indptr, int32_t num_rows, int64_t num_edges,
                       int32_t *tile_row /* [n_tiles + 1] */, void *stream);

/* ------------------------------------------------------------------------------------
 * Fused edge stage.  Replaces graph.update_all(message_func, fn.sum|mean) and its
 * autograd backward (conv.py:43-47, :63) for the elementwise
 * activations; the |E| x d edge tensor is never materialised.
 *
 * One sirgcn_edge_args describes one walk over a compressed-row structure:
 *   rows      = destinations (CSR walk: forward, backward-dQ) or sources (CSC walk: backward-dK)
 *   idx[p]    = the other endpoint of the edge stored at position p
 *   eid[p]    = original edge id of position p (only needed when `e` / `de` is given)
 * Tables (dtype `dtype`): q [*, ldq], k [*, ldk], da [*, lda], e [E, lde] (optional, may be NULL),
 * out [n_rows, ldo] and de [E, ldde] (optional; backward-dQ only).
 * dst_scale / src_scale: optional fp32 per-node coefficients (NULL = 1):
 *   sum: none; mean: dst_scale = inv_in_deg; sym: dst_scale = in_norm, src_scale = out_norm.
 * They are indexed by the destination / source node id as seen by this walk, i.e. by the
 * row id or by idx[p] depending on the direction (so callers of a partitioned graph pass
 * pointers already offset to their local row range for the "row" side).
 */
typedef struct sirgcn_edge_args {
    int32_t n_rows;
    int32_t d;              /* hidden size (columns that carry data)                     */
    int32_t dtype;          /* SIRGCN_F32 / BF16 / F16                                   */
    int32_t act;            /* SIRGCN_ACT_*                                              */
    float act_param;        /* LeakyReLU negative slope                                  */
    int32_t long_threshold; /* rows with degree > this are handled via the schedule      */
    const int32_t *indptr;  /* [n_rows + 1]                                              */
    const int32_t *idx;     /* [E]                                                       */
    const int32_t *eid;     /* [E] or NULL                                               */
    const void *q;  int64_t ldq;
    const void *k;  int64_t ldk;
    const void *da; int64_t lda;    /* backward only                                      */
    const void *e;  int64_t lde;    /* optional projected edge term, indexed by edge id   */
    void *out;      int64_t ldo;
    void *de;       int64_t ldde;   /* optional, backward-dQ only: gradient of e          */
    void *da_scaled; int64_t ldds;  /* optional, backward-dQ only: dA[u] * dst_scale[u] per row, so that the
                                       CSC walk can gather an already scaled table (then called with
                                       da = da_scaled, dst_scale = NULL); may alias `da` (in place)      */
    const float *dst_scale;
    const float *src_scale;
    /* long-row schedule of THIS walk (host-known counts; both 0 => no long rows) */
    sirgcn_schedule sched;
    int32_t n_long;
    int32_t n_chunks;
    float *partial;         /* [n_chunks, ldp] fp32 scratch, ldp = 16B-vectors*elems     */
    /* work tiles of THIS walk (sirgcn_tiles_build): tile t = rows [tile_row[t], tile_row[t+1]) */
    const int32_t *tile_row; /* [n_tiles + 1]                                            */
    int32_t n_tiles;
    int32_t accumulate;     /* != 0: out[row] += result instead of out[row] = result (a row's edges split
                               over several walks, e.g. one walk per arrived source block; fixed order)  */
    /* Edge term from a small TABLE (an nn.Embedding over edge types, benchmark-datasets/zinc/model.py:12-15): pass
     * e = the table [n_etypes, lde] and eid[p] = the type of the edge stored at position p (so no [E, d] tensor is
     * ever read).  In backward-dQ the table's gradient is reduced inside the walk: set de = NULL, n_etypes > 0 and
     * de_partial = ZEROED fp32 scratch [n_tiles + n_chunks, n_etypes, ldp] (ldp as for `partial`); every work unit
     * writes the sums of its edges per type, lane groups and units are combined in a fixed order
     * (sirgcn_etable_grad) => bitwise repeatable, no atomics.  n_etypes <= SIRGCN_MAX_ETYPES. */
    int32_t n_etypes;
    float *de_partial;
    /* SIRGCN_WALK_PLAIN_GRID: one CTA per 4 work units instead of persistent warps on a grid of the resident CTAs.
     * Persistent walks are ~6 % faster alone, but their CTAs hold every SM until the walk ends; a plain grid retires a
     * CTA every few tens of microseconds, which is what lets a collective's kernels (NCCL all-gathers of the row
     * partition) become resident WHILE a walk runs. */
    int32_t flags;
} sirgcn_edge_args;
#define SIRGCN_WALK_PLAIN_GRID 1
#define SIRGCN_MAX_ETYPES 8

/* dTable[t, c] = sum over work units u (in index order) of de_partial[u, t, c]; out is fp32 [n_etypes, ld_out]. */
int sirgcn_etable_grad(const float *de_partial, int64_t n_units, int32_t n_etypes, int32_t ldp, int32_t d,
                       float *out, int64_t ld_out, void *stream);

/* bytes of fp32 scratch needed for `partial` */
size_t sirgcn_edge_partial_bytes(int32_t n_chunks, int32_t d, int32_t dtype);

/* A[u] = dst_scale[u] * sum_{p in row u} src_scale[idx[p]] * act(q[u] + k[idx[p]] + e[eid[p]])   (CSR walk) */
int sirgcn_edge_fwd(const sirgcn_edge_args *args, void *stream);
/* dQ[u] = sum_p c_p * dA[u] (*) act'(z_p);  optionally dE[eid[p]] = that summand                  (CSR walk) */
int sirgcn_edge_bwd_q(const sirgcn_edge_args *args, void *stream);
/* dK[v] = sum_{p in row v} c_p * dA[idx[p]] (*) act'(q[idx[p]] + k[v] + e[eid[p]])                (CSC walk) */
int sirgcn_edge_bwd_k(const sirgcn_edge_args *args, void *stream);

/* ------------------------------------------------------------------------------------
 * Dense projections on the tcgen05 tensor cores (bf16 / fp16 tables, fp32 accumulation in TMEM, TMA-fed).
 * Replaces the cuBLAS GEMMs behind nn.Linear: linear_query ‖ linear_key as ONE concatenated projection
 * (conv.py:60-61), linear_relation (conv.py:65) and their input gradients.
 *   C[m, n] = A[m, k] · B[n, k]^T (+ bias[n], fp32, may be NULL)        all row-major, ld* in elements
 * forward: B = the weight as nn.Linear stores it ([out, in]);  dgrad: B = W^T (contiguous [in, out]).
 * n, k and every ld must be multiples of 8 (16-byte rows); fp32 tables are not handled here (tcgen05 has
 * no IEEE-fp32 MMA; the fp32 parity target is 1e-5). */
int sirgcn_gemm_tn(const void *a, int64_t lda, const void *b, int64_t ldb, void *c, int64_t ldc,
                   const float *bias, int64_t m, int32_t n, int32_t k, int32_t dtype, void *stream);

/* out[c] = sum over rows of x[r, c] in fp32 (bias gradients of nn.Linear: db_Q, db_R).
 * n must be a multiple of 16/sizeof(elem) and at most 256 vectors wide; two fixed-order stages. */
size_t sirgcn_colsum_workspace_bytes(int32_t n);
int sirgcn_colsum(const void *x, int64_t ld, int64_t m, int32_t n, int32_t dtype, float *out, void *workspace,
                  size_t workspace_bytes, void *stream);

/* dst[r, 0..row_bytes) = src[r, 0..row_bytes) for r < rows: row tables with different row pitches (one half of a
 * [N, 2·ld] buffer <-> a [N, ld] buffer), 128-bit vectors at HBM speed.  row_bytes, both pitches and both base
 * pointers must be multiples of 16.  Used by the memory-lean backward of the layer (no reference counterpart: the
 * reference keeps every [E, d] intermediate instead). */
int sirgcn_copy_rows(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                     int64_t row_bytes, int64_t rows, void *stream);

/* Weight and bias gradients of a projection on the tcgen05 tensor cores (16-bit tables, fp32 accumulation in TMEM):
 *     dW[n_out, k_in] = dY[m, n_out]^T · X[m, k_in]      db[n_out] = column sums of dY   (db may be NULL)
 * — the autograd backward of nn.Linear behind conv.py:60-61,:65.  Both tables are read once, as
 * stored (MN-major MMA operands fed by TMA: no transpose); the node dimension is split over the SMs and the fp32
 * partials of the splits are summed in split order (bitwise repeatable).  dW / db are fp32.  n_out, k_in, ldy, ldx
 * multiples of 8; workspace: sirgcn_gemm_wgrad_workspace_bytes(m, n_out, k_in) bytes of device scratch. */
size_t sirgcn_gemm_wgrad_workspace_bytes(int64_t m, int32_t n_out, int32_t k_in);
/* the launch plan sirgcn_gemm_wgrad would use (host-only, no CUDA call; for tests and sizing): plan[10] =
 * {node splits, node blocks (64 rows) per split, column ranges, row groups, M tiles per CTA, columns per CTA,
 *  accumulator columns per M tile, ring stages, TMEM columns, dynamic shared memory bytes} */
int sirgcn_gemm_wgrad_plan(int64_t m, int32_t n_out, int32_t k_in, int32_t with_bias, int32_t *plan);
int sirgcn_gemm_wgrad(const void *dy, int64_t ldy, const void *x, int64_t ldx, int64_t m, int32_t n_out, int32_t k_in,
                      int32_t dtype, float *dw, int64_t ld_dw, float *db, void *workspace, size_t workspace_bytes,
                      void *stream);

/* Dropout applied in place on a row table (the K / Q halves of the [N, 2·ld] projection buffer and, in backward, the
 * dK / dQ halves): table[r, c] = keep[r*d + c] ? table[r, c] * scale : 0 (fp32 product, rounded once — ATen's fused
 * dropout arithmetic), columns d..pad are zeroed.  keep = dense [rows, d] bytes (0 / 1), drawn by the host with
 * the framework's own generator in the reference's order K, Q, E (/root/reference/models/conv.py:60-61,:128). */
int sirgcn_mask_scale(void *table, int64_t pitch_bytes, const uint8_t *keep, int64_t rows, int32_t d,
                      int32_t dtype, float scale, void *stream);

/* ------------------------------------------------------------------------------------
 * Split path for arbitrary (non-elementwise) σ, agg_type 'max'/'min', and the
 * SIRConvBase/SIREConvBase classes (conv.py:47, :137-221; dictionary-lookup/model.py:17).
 * Edge tensors here ARE materialised, in CSR position order.
 */
/* z[p] = a[asel[p]] (+ b[bsel[p]]) (+ c[csel[p]]) for every edge position p < num_pos.
 * With (asel, bsel, csel) = (dst of position, src of position, edge id of position) this is the
 * pre-activation q[dst] + k[src] + e of conv.py:45/:111; with b = c = NULL it is a plain row gather
 * (edges.dst[...] / edges.src[...] of SIRConvBase, conv.py:158). No alignment requirements. */
int sirgcn_gather_add(int64_t num_pos, const int32_t *asel, const int32_t *bsel, const int32_t *csel,
                      const void *a, int64_t lda, const void *b, int64_t ldb, const void *c, int64_t ldc,
                      void *z, int64_t ldz, int32_t d, int32_t dtype, void *stream);
/* out[u] = dst_scale[u] * sum_{p in row u} src_scale[idx[p]] * m[pos(p)], summed in position order;
 * pos(p) = p when perm == NULL else perm[p] (CSC-ordered reduction of a CSR-ordered edge tensor).
 * Replaces fn.sum / fn.mean (conv.py:41,63) and the index_add_ backward of the gathers (K11). */
int sirgcn_segment_sum(int32_t n_rows, const int32_t *indptr, const int32_t *idx, const int32_t *perm,
                       const void *m, int64_t ldm, void *out, int64_t ldo, int32_t d, int32_t dtype,
                       const float *dst_scale, const float *src_scale, void *stream);
/* out[u,c] = max/min_{p in row u} m[p,c] (0 for empty rows, as DGL's fn.max), arg[u,c] = winning
 * position (first extremum in position order) or -1.  Replaces fn.max/fn.min (conv.py:41,47). */
int sirgcn_segment_minmax(int32_t n_rows, const int32_t *indptr, const void *m, int64_t ldm,
                          void *out, int64_t ldo, int32_t *arg, int64_t ldarg, int32_t d, int32_t dtype,
                          int32_t is_min, void *stream);
/* dm[p,c] = (arg[rsel[p],c] == p) ? dout[rsel[p],c] : 0, rsel[p] = row of position p */
int sirgcn_segment_minmax_bwd(int64_t num_pos, const int32_t *rsel, const void *dout, int64_t lddo,
                              const int32_t *arg, int64_t ldarg, void *dm, int64_t lddm,
                              int32_t d, int32_t dtype, void *stream);

/* ------------------------------------------------------------------------------------
 * Peer-memory transport of the row-partitioned graph (one process per GPU of a node; SURVEY.md §8e — a new
 * capability, the reference is single-process).  Every rank keeps its slice of a row table (K, Q or the scaled
 * dA) in a buffer made by sirgcn_peer_alloc; the peers map it with sirgcn_peer_open (CUDA IPC) and the all-gather
 * of the table is world-1 sirgcn_peer_copy pulls on copy engines over NVLink — no SM is taken from the edge walk
 * that runs at the same time.  These two are the ONLY entry points that allocate device memory: an IPC-exportable
 * allocation cannot come from the caller's caching allocator.  sirgcn_peer_alloc zero-fills and may synchronise
 * the device (setup time only).
 * sirgcn_peer_barrier orders the ranks on the device: pads[r] (DEVICE array of `world` device pointers) is rank
 * r's flag pad, uint32[SIRGCN_PEER_MAX_WORLD] inside a peer allocation; the kernel stores `epoch` into slot
 * `rank` of every peer's pad (release, system scope) and waits until every slot of its own pad has reached
 * `epoch` (acquire); epochs must grow by one per call on every rank.  If a peer has not arrived after
 * `timeout_ns` the kernel sets status[0] = 1 (device int32) and TRAPS: the process's context is dead and every
 * later CUDA call fails, so no kernel ever reads a table that was not gathered (nothing hangs, nothing is silent). */
#define SIRGCN_IPC_HANDLE_BYTES 64
#define SIRGCN_PEER_MAX_WORLD 32
int sirgcn_peer_alloc(size_t bytes, void **dev_ptr, void *ipc_handle /* [SIRGCN_IPC_HANDLE_BYTES] out */);
int sirgcn_peer_free(void *dev_ptr);
int sirgcn_peer_open(const void *ipc_handle, void **dev_ptr);
int sirgcn_peer_close(void *dev_ptr);
int sirgcn_peer_copy(void *dst, const void *src, size_t bytes, void *stream);
/* SM-driven fan-out: every 16-byte vector of src[0..bytes) is read once and written to each of dsts[0..n_dst)
 * (HOST array of device pointers: peer mappings and/or local buffers) by a kernel of at most n_ctas CTAs —
 * posted writes over NVLink, for when the copy engines' many-to-many rate is the limit. */
int sirgcn_peer_push(const void *src, void *const *dsts, int32_t n_dst, size_t bytes, int32_t n_ctas, void *stream);
/* The same fan-out with TMA bulk copies: one elected thread per CTA streams the slice global -> shared ring ->
 * every target (cp.async.bulk), so the transfer costs a shared-memory ring per CTA and no issue slots.
 * status[0] (device int32) is set to 2 and the kernel traps if a copy never lands (bounded wait, fatal). */
int sirgcn_peer_push_tma(const void *src, void *const *dsts, int32_t n_dst, size_t bytes, int32_t n_ctas,
                         int32_t *status, void *stream);
int sirgcn_peer_barrier(uint32_t *const *pads, int32_t world, int32_t rank, uint32_t epoch, uint64_t timeout_ns,
                        int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIRGCN_H_ */
