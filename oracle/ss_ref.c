/* CPU restatement (plain C, double accumulation) of the selective-scan forward
 * and its exact adjoint.  TEST INFRASTRUCTURE ONLY: built by oracle/Makefile
 * into oracle/libss_ref.so and used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker -- never by the product path.
 *
 * Parity status: PINNED against the golden vectors generated from the
 * reference's own selective_scan_ref + torch autograd (tests/golden/).
 *
 * Math restated from (files under /root/reference/R2GenCSR/VMamba/kernels/selective_scan):
 *   forward : test_selective_scan.py:168-234 (selective_scan_ref)
 *   softplus threshold 20 : F.softplus default == fwd_kernel_oflex.cuh:126
 *   delta groups (row d reads delta row d / (dim/delta_dim)) : selective_scan_oflex.cpp:59
 *   backward formulas cross-checked with bwd_kernel_oflex.cuh:216-259
 *
 * Layouts (all contiguous, float32):
 *   u,z,out,dout,du,dz : (batch, dim, L)       delta, ddelta : (batch, delta_dim, L)
 *   A, dA : (dim, N)   B, C, dB, dC : (batch, G, N, L)   D, dD : (dim)
 *   delta_bias, ddelta_bias : (delta_dim)      last_state : (batch, dim, N)
 * Optional pointers (D, z, delta_bias and their grads) may be NULL.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double softplus_d(double x) { return x <= 20.0 ? log1p(exp(x)) : x; }
static double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

void ss_ref_fwd(const float *u, const float *delta, const float *A, const float *B,
                const float *C, const float *D, const float *z, const float *delta_bias,
                int softplus, int batch, int dim, int L, int N, int G, int delta_dim,
                float *out /* y (+D u), before the gate */, float *out_z /* gated, may be NULL */,
                float *last_state /* may be NULL */)
{
    const int rows_per_group = dim / G, rows_per_dgroup = dim / delta_dim;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < batch; ++b) {
        for (int d = 0; d < dim; ++d) {
            const int g = d / rows_per_group, dg = d / rows_per_dgroup;
            const float *ur = u + ((size_t)b * dim + d) * L;
            const float *dr = delta + ((size_t)b * delta_dim + dg) * L;
            const float *Bg = B + ((size_t)b * G + g) * N * L;
            const float *Cg = C + ((size_t)b * G + g) * N * L;
            const double bias = delta_bias ? delta_bias[dg] : 0.0;
            const double Dv = D ? D[d] : 0.0;
            double *h = (double *)calloc(N, sizeof(double));
            for (int l = 0; l < L; ++l) {
                double dt = dr[l] + bias;
                if (softplus) dt = softplus_d(dt);
                double y = 0.0;
                for (int n = 0; n < N; ++n) {
                    h[n] = exp(dt * A[(size_t)d * N + n]) * h[n] + dt * Bg[(size_t)n * L + l] * ur[l];
                    y += h[n] * Cg[(size_t)n * L + l];
                }
                y += Dv * ur[l];
                const size_t o = ((size_t)b * dim + d) * L + l;
                out[o] = (float)y;
                if (z && out_z) { double zz = z[o]; out_z[o] = (float)(y * zz * sigmoid_d(zz)); }
            }
            if (last_state) for (int n = 0; n < N; ++n) last_state[((size_t)b * dim + d) * N + n] = (float)h[n];
            free(h);
        }
    }
}

/* dout is the gradient w.r.t. the op's returned tensor: the gated output when z
 * is given, else y.  All gradient buffers are overwritten (not accumulated). */
void ss_ref_bwd(const float *u, const float *delta, const float *A, const float *B,
                const float *C, const float *D, const float *z, const float *delta_bias,
                const float *dout, int softplus, int batch, int dim, int L, int N, int G,
                int delta_dim, float *du, float *ddelta, float *dA, float *dB, float *dC,
                float *dD, float *ddelta_bias, float *dz)
{
    const int rpg = dim / G, rpd = dim / delta_dim;
    const size_t nBC = (size_t)batch * G * N * L;
    double *dBacc = (double *)calloc(nBC, sizeof(double));
    double *dCacc = (double *)calloc(nBC, sizeof(double));
    double *ddacc = (double *)calloc((size_t)batch * delta_dim * L, sizeof(double));
    double *dAacc = (double *)calloc((size_t)batch * dim * N, sizeof(double));
    double *dDacc = (double *)calloc((size_t)batch * dim, sizeof(double));
    double *dbacc = (double *)calloc((size_t)batch * dim, sizeof(double));

    /* one task per (batch, B/C group, delta group intersection) would be finest; (b, g) with the
       rows of the group walked serially keeps dB/dC and ddelta race-free as long as a delta
       group never spans two B/C groups or vice versa is handled by the per-b serial fallback */
    const int safe = (rpg % rpd == 0) || (rpd % rpg == 0);
    const int par_g = (safe && rpd <= rpg) ? G : 1; /* if delta groups are coarser than B/C groups, go serial over g */
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int b = 0; b < batch; ++b) {
        for (int gp = 0; gp < par_g; ++gp) {
            const int d_lo = par_g == 1 ? 0 : gp * rpg, d_hi = par_g == 1 ? dim : (gp + 1) * rpg;
            double *h = (double *)malloc(sizeof(double) * (size_t)(L + 1) * N);
            double *gs = (double *)malloc(sizeof(double) * N);
            double *dts = (double *)malloc(sizeof(double) * L);
            for (int d = d_lo; d < d_hi; ++d) {
                const int g = d / rpg, dg = d / rpd;
                const size_t ro = ((size_t)b * dim + d) * L;
                const float *ur = u + ro, *dr = delta + ((size_t)b * delta_dim + dg) * L;
                const float *Bg = B + ((size_t)b * G + g) * N * L, *Cg = C + ((size_t)b * G + g) * N * L;
                double *dBg = dBacc + ((size_t)b * G + g) * N * L, *dCg = dCacc + ((size_t)b * G + g) * N * L;
                const double bias = delta_bias ? delta_bias[dg] : 0.0, Dv = D ? D[d] : 0.0;
                /* forward recompute, keeping every state h[l+1][n] (h[0] = 0) */
                for (int n = 0; n < N; ++n) h[n] = 0.0;
                for (int l = 0; l < L; ++l) {
                    double dt = dr[l] + bias;
                    if (softplus) dt = softplus_d(dt);
                    dts[l] = dt;
                    for (int n = 0; n < N; ++n)
                        h[(size_t)(l + 1) * N + n] = exp(dt * A[(size_t)d * N + n]) * h[(size_t)l * N + n]
                                                   + dt * Bg[(size_t)n * L + l] * ur[l];
                }
                for (int n = 0; n < N; ++n) gs[n] = 0.0;
                for (int l = L - 1; l >= 0; --l) {
                    const double dt = dts[l], uv = ur[l];
                    double dy = dout[ro + l];
                    if (z) {
                        double y = Dv * uv;
                        for (int n = 0; n < N; ++n) y += h[(size_t)(l + 1) * N + n] * Cg[(size_t)n * L + l];
                        const double zz = z[ro + l], s = sigmoid_d(zz);
                        if (dz) dz[ro + l] = (float)(dy * y * s * (1.0 + zz * (1.0 - s)));
                        dy *= zz * s;
                    }
                    double du_v = dy * Dv, ddt = 0.0;
                    dDacc[(size_t)b * dim + d] += dy * uv;
                    for (int n = 0; n < N; ++n) {
                        const double An = A[(size_t)d * N + n], Bn = Bg[(size_t)n * L + l], Cn = Cg[(size_t)n * L + l];
                        /* gs[n] currently holds a[l+1] * g[l+1] */
                        const double gcur = dy * Cn + gs[n];
                        const double a = exp(dt * An);
                        const double ah = a * h[(size_t)l * N + n];          /* a[l] * h[l-1] */
                        dCg[(size_t)n * L + l] += dy * h[(size_t)(l + 1) * N + n];
                        dBg[(size_t)n * L + l] += gcur * dt * uv;
                        du_v += gcur * dt * Bn;
                        ddt += gcur * Bn * uv + gcur * ah * An;
                        dAacc[((size_t)b * dim + d) * N + n] += gcur * ah * dt;
                        gs[n] = a * gcur;
                    }
                    if (softplus) { const double raw = dr[l] + bias; if (raw <= 20.0) ddt *= sigmoid_d(raw); }
                    du[ro + l] = (float)du_v;
                    ddacc[((size_t)b * delta_dim + dg) * L + l] += ddt;
                    dbacc[(size_t)b * dim + d] += ddt;
                }
            }
            free(h); free(gs); free(dts);
        }
    }
    for (size_t i = 0; i < nBC; ++i) { dB[i] = (float)dBacc[i]; dC[i] = (float)dCacc[i]; }
    for (size_t i = 0; i < (size_t)batch * delta_dim * L; ++i) ddelta[i] = (float)ddacc[i];
    for (int d = 0; d < dim; ++d) {
        for (int n = 0; n < N; ++n) {
            double s = 0.0;
            for (int b = 0; b < batch; ++b) s += dAacc[((size_t)b * dim + d) * N + n];
            dA[(size_t)d * N + n] = (float)s;
        }
        if (dD) { double s = 0.0; for (int b = 0; b < batch; ++b) s += dDacc[(size_t)b * dim + d]; dD[d] = (float)s; }
    }
    if (ddelta_bias) {
        for (int dg = 0; dg < delta_dim; ++dg) {
            double s = 0.0;
            for (int b = 0; b < batch; ++b) for (int r = 0; r < rpd; ++r) s += dbacc[(size_t)b * dim + dg * rpd + r];
            ddelta_bias[dg] = (float)s;
        }
    }
    free(dBacc); free(dCacc); free(ddacc); free(dAacc); free(dDacc); free(dbacc);
}
