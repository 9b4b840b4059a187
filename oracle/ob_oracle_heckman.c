/* oracle/ob_oracle_heckman.c -- CPU restatement of the reference's Heckman two-step replicate
 * (SURVEY.md 8f-4).  TEST INFRASTRUCTURE ONLY (see ob_oracle.h).
 *
 * Restates, citing /root/reference/oaxaca_blinder/src/:
 *   math/probit.rs:25-175      Fisher scoring from beta = 0, <= max_iter steps, ||step|| < tol; -H factored by
 *                              Cholesky, LU fallback; regulariser 1e-9 on the diagonal; Phi clamped to [1e-10, 1-1e-10]
 *   heckman.rs:38-108          probit -> IMR phi/Phi on the selected rows (0 when Phi < 1e-10) -> OLS of y on [X | IMR]
 *                              -> delta = mean(-lambda (lambda + z'gamma))
 *   estimation.rs:114-269      HeckmanEstimator: selection matrix [1 | selection predictors] over ALL rows of the group,
 *                              outcome equation over the rows with selection == 1, coefficient / mean vectors of
 *                              length K+1 (IMR last), residuals zero, no Yun normalisation, weights ignored
 *   builder.rs:464-534         detailed_selection: theta_ref * delta_ref * gamma_ref[i] * (zbar_a[i] - zbar_b[i]), ref =
 *                              group A for GroupA, group B otherwise
 *   builder.rs:538-699         beta*, two/three-fold, detailed rows over the K+1 columns, total gap over all rows
 *
 * Parity status: "parity unpinned" in the reference itself -- its tests assert only that probit converges with a
 * positive slope (probit.rs:180-211) and that a row named "IMR" exists (tests/heckman_test.rs).  The probit and the
 * two-step are therefore pinned against an independent numpy/scipy restatement (tests/golden/make_heckman_golden.py);
 * normal pdf / cdf follow statrs (pdf = exp(-x^2/2)/sqrt(2 pi), cdf = erfc(-x/sqrt 2)/2).
 * Pooled | Neumark with Heckman: the reference's pooled regression yields a K-vector beta* against K+1-vectors
 * beta_a / beta_b (builder.rs:548-589 vs estimation.rs:139-141) -- a dimension mismatch; refused here (ORC_ERR_POLARS). */
#include "ob_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;

static double norm_pdf(double x) { return exp(-0.5 * x * x) / 2.5066282746310002; }   /* statrs Normal::pdf, sqrt(2 pi) */
static double norm_cdf(double x) { return 0.5 * erfc(-x / 1.4142135623730951); }        /* statrs Normal::cdf */

/* Cholesky of an SPD matrix in place (lower), nalgebra semantics: fails on a pivot <= 0 or NaN */
static int chol_small(ld* A, int n) {
    for (int j = 0; j < n; ++j) {
        for (int k = 0; k < j; ++k) {
            const ld f = A[j * n + k];
            for (int i = j; i < n; ++i) A[i * n + j] -= A[i * n + k] * f;
        }
        const ld d = A[j * n + j];
        if (!(d > 0.0L)) return 0;
        const ld r = sqrtl(d);
        A[j * n + j] = r;
        for (int i = j + 1; i < n; ++i) A[i * n + j] /= r;
    }
    return 1;
}
static void chol_small_solve(const ld* L, int n, ld* b) {
    for (int i = 0; i < n; ++i) { ld s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { ld s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}
/* LU with partial pivoting (nalgebra lu().solve): returns 0 when a pivot is exactly zero */
static int lu_solve(ld* A, int n, ld* b) {
    for (int c = 0; c < n; ++c) {
        int piv = c; ld best = fabsl(A[c * n + c]);
        for (int r = c + 1; r < n; ++r) if (fabsl(A[r * n + c]) > best) { best = fabsl(A[r * n + c]); piv = r; }
        if (A[piv * n + c] == 0.0L) return 0;
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const ld t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
            const ld t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < n; ++r) {
            const ld f = A[r * n + c] / A[c * n + c];
            for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
            b[r] -= f * b[c];
        }
    }
    for (int i = n - 1; i >= 0; --i) { ld s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
    return 1;
}

/* math/probit.rs:25-175.  X row-major [n x k]; returns ORC_OK / ORC_ERR_NALGEBRA. */
int orc_probit(const double* y, const double* X, int64_t n, int32_t k, int32_t max_iter, double tol,
               double* beta, int32_t* converged, int32_t* iterations) {
    ld* H = (ld*)malloc(sizeof(ld) * (size_t)k * k);
    ld* g = (ld*)malloc(sizeof(ld) * (size_t)k);
    ld* step = (ld*)malloc(sizeof(ld) * (size_t)k);
    ld* M = (ld*)malloc(sizeof(ld) * (size_t)k * k);
    for (int j = 0; j < k; ++j) beta[j] = 0.0;                                  /* :41 */
    *converged = 0; *iterations = 0;
    int rc = ORC_OK;
    for (int it = 0; it < max_iter; ++it) {                                      /* :50 */
        *iterations = it + 1;
        for (int j = 0; j < k * k; ++j) H[j] = 0.0L;
        for (int j = 0; j < k; ++j) g[j] = 0.0L;
        for (int64_t i = 0; i < n; ++i) {
            const double* x = X + i * k;
            double z = 0.0;
            for (int j = 0; j < k; ++j) z += x[j] * beta[j];                     /* :54 */
            const double phi = norm_pdf(z);
            double Phi = norm_cdf(z);
            if (Phi < 1e-10) Phi = 1e-10;                                        /* :71 clamp */
            if (Phi > 1.0 - 1e-10) Phi = 1.0 - 1e-10;
            const double lam = y[i] > 0.5 ? phi / Phi : -phi / (1.0 - Phi);       /* :73-77 */
            const double sw = sqrt((phi * phi) / (Phi * (1.0 - Phi)));           /* :82-83 */
            const double w = sw * sw;                                            /* :95-98: weight = sqrt(w)^2 */
            for (int j = 0; j < k; ++j) {
                g[j] += (ld)lam * x[j];                                          /* :87 */
                for (int l = 0; l <= j; ++l) H[j * k + l] -= (ld)x[j] * x[l] * w; /* :100-116 */
            }
        }
        for (int j = 0; j < k; ++j) for (int l = j + 1; l < k; ++l) H[j * k + l] = H[l * k + j];
        for (int j = 0; j < k; ++j) H[j * k + j] -= 1e-9L;                       /* :124-126 */
        for (int j = 0; j < k * k; ++j) M[j] = -H[j];                            /* :132 */
        for (int j = 0; j < k; ++j) step[j] = g[j];
        if (chol_small(M, k)) chol_small_solve(M, k, step);                      /* :133-134 */
        else {                                                                   /* :135-147 LU fallback: H s = g, step = -s */
            for (int j = 0; j < k * k; ++j) M[j] = H[j];
            for (int j = 0; j < k; ++j) step[j] = g[j];
            if (!lu_solve(M, k, step)) { rc = ORC_ERR_NALGEBRA; break; }
            for (int j = 0; j < k; ++j) step[j] = -step[j];
        }
        ld nrm = 0.0L;
        for (int j = 0; j < k; ++j) { beta[j] = (double)((ld)beta[j] + step[j]); nrm += step[j] * step[j]; }   /* :150 */
        if (sqrtl(nrm) < (ld)tol) { *converged = 1; break; }                     /* :152-155 */
    }
    if (rc == ORC_OK) {   /* :166-173 vcov = -(H^-1): fails only on an exactly singular H */
        for (int j = 0; j < k * k; ++j) M[j] = H[j];
        for (int j = 0; j < k; ++j) step[j] = 0.0L;
        if (!lu_solve(M, k, step)) rc = ORC_ERR_NALGEBRA;
    }
    free(H); free(g); free(step); free(M);
    return rc;
}

/* heckman.rs:38-108 on one (resampled) group.  X [n x K] outcome design, Z [n x K1] selection design (intercept
 * first), s [n] selection outcome.  Outputs: beta_aug [K+1] (IMR coefficient last), x_mean_aug [K+1], gamma [K1],
 * z_mean [K1] (over ALL rows, estimation.rs:172), delta. */
static int heckman_group(const double* X, const double* y, const double* Z, const double* s, int64_t n, int K, int K1,
                         int precise, double* beta_aug, double* x_mean_aug, double* gamma, double* z_mean, double* delta) {
    int32_t conv = 0, iters = 0;
    int rc = orc_probit(s, Z, n, K1, 100, 1e-6, gamma, &conv, &iters);             /* heckman.rs:46 */
    if (rc != ORC_OK) return rc;
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) m += (s[i] == 1.0);                            /* estimation.rs:213 equal(1) */
    if (m == 0) return ORC_ERR_INVALID_GROUP;                                       /* estimation.rs:236-240 */
    const int Ka = K + 1;
    double* Xa = (double*)malloc(sizeof(double) * (size_t)m * Ka);
    double* ya = (double*)malloc(sizeof(double) * (size_t)m);
    ld imr_sum = 0.0L, delta_sum = 0.0L;
    int64_t r = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (s[i] != 1.0) continue;
        double zg = 0.0;
        for (int j = 0; j < K1; ++j) zg += Z[i * K1 + j] * gamma[j];               /* heckman.rs:54 */
        const double phi = norm_pdf(zg), Phi = norm_cdf(zg);
        const double imr = Phi < 1e-10 ? 0.0 : phi / Phi;                          /* :58-66 */
        memcpy(Xa + r * Ka, X + i * K, sizeof(double) * (size_t)K);
        Xa[r * Ka + K] = imr;                                                      /* :72-75 */
        ya[r] = y[i];
        imr_sum += imr;
        delta_sum += -imr * (imr + zg);                                            /* :92-97 */
        ++r;
    }
    rc = orc_ols(ya, Xa, NULL, m, Ka, precise, beta_aug, NULL);                    /* :78 */
    if (rc == ORC_OK) {
        for (int j = 0; j < K; ++j) {                                              /* estimation.rs:143-144 row_mean of the filtered X */
            ld a = 0.0L;
            for (int64_t q = 0; q < m; ++q) a += Xa[q * Ka + j];
            x_mean_aug[j] = (double)(a / (ld)m);
        }
        x_mean_aug[K] = (double)(imr_sum / (ld)m);                                 /* :146-151 */
        *delta = (double)(delta_sum / (ld)m);
        for (int j = 0; j < K1; ++j) {                                             /* :171-172 */
            ld a = 0.0L;
            for (int64_t i = 0; i < n; ++i) a += Z[i * K1 + j];
            z_mean[j] = (double)(a / (ld)n);
        }
    }
    free(Xa); free(ya);
    return rc;
}

int32_t orc_heckman_n_stats(int32_t K, int32_t K1) { return 5 + 2 * (K + 1) + K1; }

/* One pass (builder.rs:420-699 with the HeckmanEstimator).  stats layout:
 * [explained, unexplained, endowments, coefficients, interaction, det_expl[K+1], det_unexpl[K+1], selection[K1]] */
int orc_heckman_pass(int32_t K, int32_t K1, int32_t ref_kind,
                     const double* Xa, const double* ya, const double* Za, const double* sa, int64_t na,
                     const double* Xb, const double* yb, const double* Zb, const double* sb, int64_t nb,
                     int precise, double* stats, double* beta_a, double* beta_b, double* gamma_a, double* gamma_b,
                     double* total_gap) {
    if (na == 0 || nb == 0) return ORC_ERR_INVALID_GROUP;
    if (ref_kind == ORC_REF_POOLED) return ORC_ERR_POLARS;
    const int Ka = K + 1;
    double* xa = (double*)malloc(sizeof(double) * (size_t)Ka); double* xb = (double*)malloc(sizeof(double) * (size_t)Ka);
    double* za = (double*)malloc(sizeof(double) * (size_t)K1); double* zb = (double*)malloc(sizeof(double) * (size_t)K1);
    double* bs = (double*)malloc(sizeof(double) * (size_t)Ka);
    double da = 0.0, db = 0.0;
    int rc = heckman_group(Xa, ya, Za, sa, na, K, K1, precise, beta_a, xa, gamma_a, za, &da);   /* estimation.rs:132 */
    if (rc == ORC_OK) rc = heckman_group(Xb, yb, Zb, sb, nb, K, K1, precise, beta_b, xb, gamma_b, zb, &db);
    if (rc == ORC_OK) {
        const int D = Ka;
        /* builder.rs:477-534 */
        const int ref_a = ref_kind == ORC_REF_GROUP_A;
        const double theta = ref_a ? beta_a[K] : beta_b[K], del = ref_a ? da : db;
        const double* gam = ref_a ? gamma_a : gamma_b;
        for (int j = 0; j < K1; ++j) stats[5 + 2 * D + j] = theta * del * gam[j] * (za[j] - zb[j]);
        /* builder.rs:538-621 */
        if (ref_kind == ORC_REF_GROUP_A) memcpy(bs, beta_a, sizeof(double) * (size_t)Ka);
        else if (ref_kind == ORC_REF_GROUP_B) memcpy(bs, beta_b, sizeof(double) * (size_t)Ka);
        else {
            const double total = (double)na + (double)nb;             /* df_a.height(), df_b.height(): all rows (no weights here) */
            const double wA = (double)na / total, wB = 1.0 - wA;
            for (int j = 0; j < Ka; ++j) bs[j] = beta_a[j] * wA + beta_b[j] * wB;
        }
        double two[2], three[3];
        orc_three_fold(xa, xb, beta_a, beta_b, Ka, three);
        orc_two_fold(xa, xb, beta_a, beta_b, bs, Ka, two);
        orc_detailed(xa, xb, beta_a, beta_b, bs, Ka, stats + 5, stats + 5 + D);
        stats[0] = two[0]; stats[1] = two[1]; stats[2] = three[0]; stats[3] = three[1]; stats[4] = three[2];
        ld ma = 0.0L, mb = 0.0L;                                       /* builder.rs:676-684: y over ALL rows of each group */
        for (int64_t i = 0; i < na; ++i) ma += ya[i];
        for (int64_t i = 0; i < nb; ++i) mb += yb[i];
        *total_gap = (double)(ma / (ld)na - mb / (ld)nb);
    }
    free(xa); free(xb); free(za); free(zb); free(bs);
    return rc;
}

/* builder.rs:787-951 with the Heckman estimator: point pass + replicates under an explicit index stream
 * (idx_a [reps x na], idx_b [reps x nb]) + bootstrap_stats over the successful replicates. */
int orc_heckman_run(int32_t K, int32_t K1, int32_t ref_kind,
                    const double* Xa, const double* ya, const double* Za, const double* sa, int64_t na,
                    const double* Xb, const double* yb, const double* Zb, const double* sb, int64_t nb,
                    int64_t reps, const uint32_t* idx_a, const uint32_t* idx_b, int nthreads, int precise,
                    double* point_stats, double* point_beta_a, double* point_beta_b, double* point_gamma_a, double* point_gamma_b,
                    double* total_gap, double* rep_stats, int32_t* rep_status, double* rep_gamma_a,
                    int64_t* n_ok, double* se, double* p, double* ci_lo, double* ci_hi, double* t) {
    const int S = orc_heckman_n_stats(K, K1), Ka = K + 1;
    int rc = orc_heckman_pass(K, K1, ref_kind, Xa, ya, Za, sa, na, Xb, yb, Zb, sb, nb, precise, point_stats, point_beta_a,
                              point_beta_b, point_gamma_a, point_gamma_b, total_gap);
    if (rc != ORC_OK) return rc;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < reps; ++b) {
        double* gXa = (double*)malloc(sizeof(double) * (size_t)na * K); double* gya = (double*)malloc(sizeof(double) * (size_t)na);
        double* gZa = (double*)malloc(sizeof(double) * (size_t)na * K1); double* gsa = (double*)malloc(sizeof(double) * (size_t)na);
        double* gXb = (double*)malloc(sizeof(double) * (size_t)nb * K); double* gyb = (double*)malloc(sizeof(double) * (size_t)nb);
        double* gZb = (double*)malloc(sizeof(double) * (size_t)nb * K1); double* gsb = (double*)malloc(sizeof(double) * (size_t)nb);
        for (int64_t i = 0; i < na; ++i) {   /* sample_n_literal gathers every column (builder.rs:822-827) */
            const int64_t r = idx_a[b * na + i];
            memcpy(gXa + i * K, Xa + r * K, sizeof(double) * (size_t)K); memcpy(gZa + i * K1, Za + r * K1, sizeof(double) * (size_t)K1);
            gya[i] = ya[r]; gsa[i] = sa[r];
        }
        for (int64_t i = 0; i < nb; ++i) {
            const int64_t r = idx_b[b * nb + i];
            memcpy(gXb + i * K, Xb + r * K, sizeof(double) * (size_t)K); memcpy(gZb + i * K1, Zb + r * K1, sizeof(double) * (size_t)K1);
            gyb[i] = yb[r]; gsb[i] = sb[r];
        }
        double* ba = (double*)malloc(sizeof(double) * (size_t)Ka); double* bb = (double*)malloc(sizeof(double) * (size_t)Ka);
        double* ga = (double*)malloc(sizeof(double) * (size_t)K1); double* gb = (double*)malloc(sizeof(double) * (size_t)K1);
        double gap = 0.0;
        const int st = orc_heckman_pass(K, K1, ref_kind, gXa, gya, gZa, gsa, na, gXb, gyb, gZb, gsb, nb, precise,
                                        rep_stats + b * S, ba, bb, ga, gb, &gap);
        rep_status[b] = st;
        if (st != ORC_OK) for (int j = 0; j < S; ++j) rep_stats[b * S + j] = NAN;
        if (rep_gamma_a) for (int j = 0; j < K1; ++j) rep_gamma_a[b * K1 + j] = st == ORC_OK ? ga[j] : NAN;
        free(gXa); free(gya); free(gZa); free(gsa); free(gXb); free(gyb); free(gZb); free(gsb);
        free(ba); free(bb); free(ga); free(gb);
    }
    orc_reduce(rep_stats, rep_status, reps, S, point_stats, n_ok, se, p, ci_lo, ci_hi, t);
    return ORC_OK;
}
