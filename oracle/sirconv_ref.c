/* Plain-C restatement of the SIR-GCN edge stage and its analytic backward — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A second, independent statement of what /root/reference/models/conv.py computes between the projections
 * (message_func :43-47 / :109-113 and update_all with fn.sum / fn.mean :63 / :130), written as explicit loops over
 * the COO edge list in double precision, with the derivative written out by hand instead of obtained from autograd.
 * tests/test_oracle_pins.py checks it against oracle/sirconv_ref.py (torch ops + autograd): two restatements that
 * share no code must agree to 1e-12 (1e-6 for 'sym': torch.pow(x,-0.5) vs IEEE 1/sqrt(x) in fp32).  The torch restatement
 * is itself pinned on the reference's own code executed through tests/fake_dgl (see its header); DGL's own arithmetic
 * stays unpinned (DGL is not installable here).
 *
 *   z_e  = q[dst_e] + k[src_e] (+ e_e)                                   conv.py:45 / :111
 *   m_e  = c_e * act(z_e),  c_e = out_norm[src_e] * in_norm[dst_e]       conv.py:45-46, :51-57 ('sym'; 1 otherwise)
 *   A[u] = sum over in-edges of m_e           ('sum', 'sym')             conv.py:63  fn.sum
 *        = that sum / max(in_deg(u), 1)       ('mean')                   conv.py:41  fn.mean
 *   backward, given dA:   g_e = c'_e * dA[dst_e] * act'(z_e)   (c' includes the 1/deg of 'mean')
 *                         dQ[u] = sum_{e: dst_e = u} g_e,  dK[v] = sum_{e: src_e = v} g_e,  dE_e = g_e
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_LEAKY_RELU = 2, ACT_GELU = 3 };
enum { AGG_SUM = 0, AGG_MEAN = 1, AGG_SYM = 2 };

static double act_f(double z, int act, double p) {
    switch (act) {
        case ACT_RELU: return z > 0 ? z : 0;
        case ACT_LEAKY_RELU: return z > 0 ? z : p * z;
        case ACT_GELU: return 0.5 * z * (1.0 + erf(z * 0.70710678118654752440));      /* nn.GELU() default (erf) */
        default: return z;
    }
}

static double act_d(double z, int act, double p) {       /* ATen: relu'(0) = 0, leaky_relu'(0) = slope */
    switch (act) {
        case ACT_RELU: return z > 0 ? 1 : 0;
        case ACT_LEAKY_RELU: return z > 0 ? 1 : p;
        case ACT_GELU: return 0.5 * (1.0 + erf(z * 0.70710678118654752440)) +
                              z * 0.39894228040143267794 * exp(-0.5 * z * z);
        default: return 1;
    }
}

/* per-edge coefficient c_e (and the 1/deg of 'mean', which fn.mean applies after the sum: same product) */
static int coefficients(int64_t E, int32_t N, const int64_t *src, const int64_t *dst, int agg, double **coef_out) {
    double *in_deg = calloc((size_t)N > 0 ? N : 1, sizeof(double)), *out_deg = calloc((size_t)N > 0 ? N : 1, sizeof(double));
    double *coef = malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    if (!in_deg || !out_deg || !coef) return -2;
    for (int64_t e = 0; e < E; ++e) {
        if (src[e] < 0 || src[e] >= N || dst[e] < 0 || dst[e] >= N) return -1;
        in_deg[dst[e]] += 1;
        out_deg[src[e]] += 1;
    }
    for (int64_t e = 0; e < E; ++e) {
        const double di = in_deg[dst[e]] < 1 ? 1 : in_deg[dst[e]], dq = out_deg[src[e]] < 1 ? 1 : out_deg[src[e]];
        /* the reference takes the norms in fp32 (conv.py:51-52 `.float()`, then torch.pow(., -0.5)): fp32 values here too
         * (IEEE 1/sqrt; torch's pow may differ from it in the last fp32 bit) */
        coef[e] = agg == AGG_SYM ? (double)(1.0f / sqrtf((float)dq)) * (double)(1.0f / sqrtf((float)di)) : agg == AGG_MEAN ? 1.0 / di : 1.0;
    }
    free(in_deg);
    free(out_deg);
    *coef_out = coef;
    return 0;
}

int sirconv_ref_edge_forward(int64_t E, int32_t N, int32_t d, const int64_t *src, const int64_t *dst,
                             const double *q, const double *k, const double *e /* [E,d] or NULL */,
                             int act, double slope, int agg, double *a_out /* [N,d] */) {
    double *coef;
    int rc = coefficients(E, N, src, dst, agg, &coef);
    if (rc) return rc;
    memset(a_out, 0, sizeof(double) * (size_t)N * d);
    for (int64_t i = 0; i < E; ++i)
        for (int c = 0; c < d; ++c) {
            const double z = q[dst[i] * d + c] + k[src[i] * d + c] + (e ? e[i * d + c] : 0.0);
            a_out[dst[i] * d + c] += coef[i] * act_f(z, act, slope);
        }
    free(coef);
    return 0;
}

int sirconv_ref_edge_backward(int64_t E, int32_t N, int32_t d, const int64_t *src, const int64_t *dst,
                              const double *q, const double *k, const double *e, const double *da /* [N,d] */,
                              int act, double slope, int agg,
                              double *dq /* [N,d] */, double *dk /* [N,d] */, double *de /* [E,d] or NULL */) {
    double *coef;
    int rc = coefficients(E, N, src, dst, agg, &coef);
    if (rc) return rc;
    memset(dq, 0, sizeof(double) * (size_t)N * d);
    memset(dk, 0, sizeof(double) * (size_t)N * d);
    for (int64_t i = 0; i < E; ++i)
        for (int c = 0; c < d; ++c) {
            const double z = q[dst[i] * d + c] + k[src[i] * d + c] + (e ? e[i * d + c] : 0.0);
            const double g = coef[i] * da[dst[i] * d + c] * act_d(z, act, slope);
            dq[dst[i] * d + c] += g;
            dk[src[i] * d + c] += g;
            if (de) de[i * d + c] = g;
        }
    free(coef);
    return 0;
}
