/* CPU oracle for graph index construction — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates what the reference obtains from DGL when the layer asks for degrees and
 * runs update_all over the in-CSR (/root/reference/models/conv.py:51-52, :63):
 * a destination-major ordering of the COO edge list in which ties keep edge-id
 * order (DGL's COO->CSR conversion is a stable counting sort), the mirror
 * source-major ordering used by the backward scatter (index_add_ over src,
 * SURVEY.md K11), and the clamp(min=1)^-1/2 norms of conv.py:51-57.
 *
 * PARITY UNPINNED against DGL for this file (DGL 2.1.0, requirements.txt:1, is not installed and
 * the reference has no fixtures: the stable tie order is DGL's documented behaviour, not a value the
 * reference's own code produces); pinned against torch.sort(stable=True) in tests/test_oracle_pins.py.
 *
 * Plain C, single thread, O(N+E) counting sort.  Built by oracle/Makefile into
 * oracle/_build/libcsr_ref.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int counting_sort(const int32_t *key, const int32_t *other, int64_t E, int32_t N,
                         int32_t *indptr, int32_t *other_sorted, int32_t *eid_sorted) {
    int64_t *cursor = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    if (!cursor) return -1;
    for (int64_t e = 0; e < E; ++e) {
        if (key[e] < 0 || key[e] >= N) { free(cursor); return -2; }
        cursor[key[e] + 1]++;
    }
    for (int32_t v = 0; v < N; ++v) cursor[v + 1] += cursor[v];
    for (int32_t v = 0; v <= N; ++v) indptr[v] = (int32_t)cursor[v];
    for (int64_t e = 0; e < E; ++e) {           /* ascending e => stable */
        int64_t p = cursor[key[e]]++;
        other_sorted[p] = other[e];
        eid_sorted[p] = (int32_t)e;
    }
    free(cursor);
    return 0;
}

int csr_ref_build(const int32_t *src, const int32_t *dst, int64_t E, int32_t N,
                  int32_t *indptr_in, int32_t *col_src, int32_t *eid_in,
                  int32_t *indptr_out, int32_t *row_dst, int32_t *eid_out,
                  float *in_norm, float *out_norm) {
    int rc = counting_sort(dst, src, E, N, indptr_in, col_src, eid_in);
    if (rc) return rc;
    rc = counting_sort(src, dst, E, N, indptr_out, row_dst, eid_out);
    if (rc) return rc;
    for (int32_t v = 0; v < N; ++v) {
        int32_t di = indptr_in[v + 1] - indptr_in[v];
        int32_t dq = indptr_out[v + 1] - indptr_out[v];
        in_norm[v] = 1.0f / sqrtf((float)(di < 1 ? 1 : di));
        out_norm[v] = 1.0f / sqrtf((float)(dq < 1 ? 1 : dq));
    }
    return 0;
}
