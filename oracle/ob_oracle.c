/* oracle/ob_oracle.c -- CPU restatement of the reference's bootstrap-inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ob_oracle.h).  Plain C99 + OpenMP, no dependencies.
 * Citations are file:line under /root/reference/oaxaca_blinder/src/.
 *
 * Two arithmetic modes:
 *   precise = 1  X'WX and X'Wy are accumulated in double-double (exact products, compensated sums;
 *                ~106 bits) -> the oracle is the more accurate side of every GPU-vs-oracle
 *                comparison (checker mode).
 *   precise = 0  plain double, same algorithmic steps as the reference incl. the work it throws
 *                away per replicate (y_hat, residuals, (X'X)^-1; ols.rs:118-137) -> the CPU
 *                baseline that bench.py times.
 * All K x K work (Cholesky, solves) is done in long double in both modes: it is O(K^3) noise.
 */
#include "ob_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;

int32_t orc_n_base(const orc_spec* s) {
    int32_t nb = 0;
    for (int32_t v = 0; v < s->n_norm; ++v) nb += s->norm_has_base[v] ? 1 : 0;
    return nb;
}

int32_t orc_n_stats(const orc_spec* s) { return 5 + 2 * (s->K + orc_n_base(s)); }

/* ---- nalgebra Cholesky semantics (nalgebra 0.32 linalg/cholesky.rs, restated from its
 * published algorithm: column-by-column; a pivot that is zero, negative or NaN -> None).
 * Call site ols.rs:107-111.  G is K x K row-major, lower triangle is overwritten by L. ---- */
/* smallest pivot relative to its original diagonal seen since the last reset: 1 - R^2 of the most
 * collinear column.  Tests use it to recognise numerically rank-deficient resamples, where the sign of
 * a rounding-noise pivot -- and hence success/failure -- is not reproducible across summation orders. */
static _Thread_local double tls_min_pivot = 1.0;

static int chol_factor(ld* G, int K) {
    ld* diag0 = (ld*)malloc(sizeof(ld) * (size_t)K);
    for (int j = 0; j < K; ++j) diag0[j] = G[j * K + j];
    for (int j = 0; j < K; ++j) {
        for (int k = 0; k < j; ++k) {
            const ld f = G[j * K + k];
            for (int i = j; i < K; ++i) G[i * K + j] -= G[i * K + k] * f;
        }
        const ld d = G[j * K + j];
        if (diag0[j] > 0.0L) { /* an exactly zero column (absent category) is singular in any arithmetic: not ambiguous */
            const double rel = (double)(d / diag0[j]);
            if (!(rel >= tls_min_pivot)) tls_min_pivot = rel;
        }
        if (!(d > 0.0L)) { free(diag0); return 0; }
        const ld r = sqrtl(d);
        G[j * K + j] = r;
        for (int i = j + 1; i < K; ++i) G[i * K + j] /= r;
    }
    free(diag0);
    return 1;
}

double orc_last_min_pivot(void) { return tls_min_pivot; }
void orc_reset_min_pivot(void) { tls_min_pivot = 1.0; }

static void chol_solve(const ld* L, int K, ld* b) { /* ols.rs:115 */
    for (int i = 0; i < K; ++i) {
        ld s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * K + k] * b[k];
        b[i] = s / L[i * K + i];
    }
    for (int i = K - 1; i >= 0; --i) {
        ld s = b[i];
        for (int k = i + 1; k < K; ++k) s -= L[k * K + i] * b[k];
        b[i] = s / L[i * K + i];
    }
}

/* (X'X)^-1 from the factor: ols.rs:136.  Only the baseline mode pays for it (the result is
 * never read on this path, exactly as in the reference). */
static double chol_inverse_trace(const ld* L, int K) {
    ld* col = (ld*)malloc(sizeof(ld) * (size_t)K);
    ld tr = 0;
    for (int j = 0; j < K; ++j) {
        for (int i = 0; i < K; ++i) col[i] = (i == j) ? 1.0L : 0.0L;
        chol_solve(L, K, col);
        tr += col[j];
    }
    free(col);
    return (double)tr;
}

/* ---- baseline-mode X'X and X'y (ols.rs:80-81, :88-89) --------------------------------------
 * nalgebra's `x.transpose() * x` dispatches to the matrixmultiply crate: a cache-blocked,
 * register-tiled SIMD dgemm (packed panels, 4 x 12-ish micro-kernel on AVX2/FMA), computing the
 * FULL K x K product -- it does not know the result is symmetric.  This restates that algorithm
 * shape: row blocks of X are packed into a zero-padded panel, and a 4 x 12 FMA micro-kernel
 * (12 ymm accumulators) sweeps every (j-block, l-block) of the full product.  Plain double
 * throughout, like the reference. */
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define GB_KC 128
static void gram_blocked(const double* X, const double* y, int64_t n, int K, double* G, double* r) {
    const int Kj = (K + 3) / 4 * 4, Kl = (K + 11) / 12 * 12, ld = Kl;
    double* P = (double*)aligned_alloc(64, sizeof(double) * (size_t)GB_KC * ld);
    double* Gp = (double*)aligned_alloc(64, sizeof(double) * (size_t)Kj * Kl);
    memset(Gp, 0, sizeof(double) * (size_t)Kj * Kl);
    memset(P, 0, sizeof(double) * (size_t)GB_KC * ld);
    for (int64_t i0 = 0; i0 < n; i0 += GB_KC) {
        const int kc = (int)((n - i0 < GB_KC) ? n - i0 : GB_KC);
        for (int k = 0; k < kc; ++k) memcpy(P + (size_t)k * ld, X + (i0 + k) * K, sizeof(double) * (size_t)K);   /* pack */
        for (int jb = 0; jb < Kj; jb += 4) {
            for (int lb = 0; lb < Kl; lb += 12) {
                __m256d c00 = _mm256_setzero_pd(), c01 = c00, c02 = c00, c10 = c00, c11 = c00, c12 = c00;
                __m256d c20 = c00, c21 = c00, c22 = c00, c30 = c00, c31 = c00, c32 = c00;
                const double* pk = P;
                for (int k = 0; k < kc; ++k, pk += ld) {
                    const __m256d b0 = _mm256_load_pd(pk + lb), b1 = _mm256_load_pd(pk + lb + 4), b2 = _mm256_load_pd(pk + lb + 8);
                    __m256d a = _mm256_broadcast_sd(pk + jb);
                    c00 = _mm256_fmadd_pd(a, b0, c00); c01 = _mm256_fmadd_pd(a, b1, c01); c02 = _mm256_fmadd_pd(a, b2, c02);
                    a = _mm256_broadcast_sd(pk + jb + 1);
                    c10 = _mm256_fmadd_pd(a, b0, c10); c11 = _mm256_fmadd_pd(a, b1, c11); c12 = _mm256_fmadd_pd(a, b2, c12);
                    a = _mm256_broadcast_sd(pk + jb + 2);
                    c20 = _mm256_fmadd_pd(a, b0, c20); c21 = _mm256_fmadd_pd(a, b1, c21); c22 = _mm256_fmadd_pd(a, b2, c22);
                    a = _mm256_broadcast_sd(pk + jb + 3);
                    c30 = _mm256_fmadd_pd(a, b0, c30); c31 = _mm256_fmadd_pd(a, b1, c31); c32 = _mm256_fmadd_pd(a, b2, c32);
                }
                double* g = Gp + (size_t)jb * Kl + lb;
#define GB_ACC(row, v0, v1, v2)                                                                  \
    _mm256_store_pd(g + (row) * Kl, _mm256_add_pd(_mm256_load_pd(g + (row) * Kl), v0));            \
    _mm256_store_pd(g + (row) * Kl + 4, _mm256_add_pd(_mm256_load_pd(g + (row) * Kl + 4), v1));    \
    _mm256_store_pd(g + (row) * Kl + 8, _mm256_add_pd(_mm256_load_pd(g + (row) * Kl + 8), v2));
                GB_ACC(0, c00, c01, c02) GB_ACC(1, c10, c11, c12) GB_ACC(2, c20, c21, c22) GB_ACC(3, c30, c31, c32)
#undef GB_ACC
            }
        }
        /* X'y: gemv over the packed block */
        for (int k = 0; k < kc; ++k) {
            const double yk = y[i0 + k];
            const double* pk = P + (size_t)k * ld;
            for (int j = 0; j < K; ++j) r[j] += pk[j] * yk;
        }
    }
    for (int j = 0; j < K; ++j)
        for (int l = 0; l < K; ++l) G[j * K + l] = Gp[(size_t)j * Kl + l];
    free(P); free(Gp);
}
#else
static void gram_blocked(const double* X, const double* y, int64_t n, int K, double* G, double* r) {
    for (int64_t i = 0; i < n; ++i) {
        const double* xi = X + i * K;
        for (int j = 0; j < K; ++j) {
            const double a = xi[j];
            for (int l = 0; l < K; ++l) G[j * K + l] += a * xi[l];
            r[j] += a * y[i];
        }
    }
}
#endif

int orc_ols(const double* y, const double* X, const double* w, int64_t n, int32_t K,
            int precise, double* beta, double* resid) {
    if (w) { /* ols.rs:60-66 */
        for (int64_t i = 0; i < n; ++i)
            if (w[i] < 0.0) return ORC_ERR_INVALID_GROUP;
    }
    ld* G = (ld*)calloc((size_t)K * K, sizeof(ld));
    ld* r = (ld*)calloc((size_t)K, sizeof(ld));
    if (precise) {
        /* ols.rs:68-81 / :88-89 with the products formed as in the reference ((sqrt(w) x_j)(sqrt(w) x_l), each factor
         * rounded to double first) but every product taken exactly (FMA two-product) and summed in double-double
         * (two-sum): ~106 significant bits, i.e. more than long double, and vectorisable -- a 10M-row pass takes seconds
         * instead of minutes, so the full-size parity tests can run the checker on the BASELINE shapes. */
        const int Kp = (K + 3) / 4 * 4;
        double* hi = (double*)aligned_alloc(64, sizeof(double) * (size_t)(K + 1) * Kp);
        double* lo = (double*)aligned_alloc(64, sizeof(double) * (size_t)(K + 1) * Kp);
        double* xr = (double*)aligned_alloc(64, sizeof(double) * (size_t)Kp);
        memset(hi, 0, sizeof(double) * (size_t)(K + 1) * Kp);
        memset(lo, 0, sizeof(double) * (size_t)(K + 1) * Kp);
        memset(xr, 0, sizeof(double) * (size_t)Kp);
        for (int64_t i = 0; i < n; ++i) {
            const double* xi = X + i * K;
            const double sw = w ? sqrt(w[i]) : 1.0;
            for (int j = 0; j < K; ++j) xr[j] = w ? xi[j] * sw : xi[j];
            const double yw = w ? y[i] * sw : y[i];
            for (int j = 0; j <= K; ++j) {      /* row K: the X'y products */
                const double a = j < K ? xr[j] : yw;
                double* restrict h = hi + (size_t)j * Kp;
                double* restrict q = lo + (size_t)j * Kp;
                const int l0 = j < K ? (j & ~3) : 0;   /* upper triangle, start rounded down to the vector width */
                for (int l = l0; l < Kp; ++l) {
                    const double b = xr[l];
                    const double p = a * b;
                    const double e = fma(a, b, -p);
                    const double t = h[l] + p;
                    const double bb = t - h[l];
                    const double err = (h[l] - (t - bb)) + (p - bb);
                    h[l] = t;
                    q[l] += err + e;
                }
            }
        }
        for (int j = 0; j < K; ++j) {
            for (int l = j; l < K; ++l) G[j * K + l] = (ld)hi[(size_t)j * Kp + l] + (ld)lo[(size_t)j * Kp + l];
            r[j] = (ld)hi[(size_t)K * Kp + j] + (ld)lo[(size_t)K * Kp + j];
        }
        free(hi); free(lo); free(xr);
        for (int j = 0; j < K; ++j)
            for (int l = 0; l < j; ++l) G[j * K + l] = G[l * K + j];
    } else {
        /* reference-shaped: materialise sqrt(w)-scaled copies (ols.rs:68-78), full K x K product */
        double* Gd = (double*)calloc((size_t)K * K, sizeof(double));
        double* rd = (double*)calloc((size_t)K, sizeof(double));
        const double* Xs = X;
        const double* ys = y;
        double* Xw = NULL;
        double* yw = NULL;
        if (w) {
            Xw = (double*)malloc(sizeof(double) * (size_t)n * K);
            yw = (double*)malloc(sizeof(double) * (size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                const double sw = sqrt(w[i]);
                for (int j = 0; j < K; ++j) Xw[i * K + j] = X[i * K + j] * sw;
                yw[i] = y[i] * sw;
            }
            Xs = Xw;
            ys = yw;
        }
        gram_blocked(Xs, ys, n, K, Gd, rd);   /* ols.rs:80-81 / :88-89: full K x K product, as nalgebra's gemm computes it */
        for (int j = 0; j < K * K; ++j) G[j] = Gd[j];
        for (int j = 0; j < K; ++j) r[j] = rd[j];
        free(Gd); free(rd); free(Xw); free(yw);
    }
    /* ols.rs:96-105: n_obs = nrows (not sum of weights) must exceed k */
    if ((double)n <= (double)K) { free(G); free(r); return ORC_ERR_INSUFFICIENT_DATA; }
    if (!chol_factor(G, K)) { free(G); free(r); return ORC_ERR_NALGEBRA; }
    chol_solve(G, K, r);
    for (int j = 0; j < K; ++j) beta[j] = (double)r[j];
    if (resid || !precise) { /* ols.rs:118-119 raw residuals y - X beta */
        double sse = 0.0;
        if (precise) {
            for (int64_t i = 0; i < n; ++i) {
                const double* xi = X + i * K;
                ld yh = 0;
                for (int j = 0; j < K; ++j) yh += (ld)xi[j] * (ld)beta[j];
                const double e = (double)((ld)y[i] - yh);
                if (resid) resid[i] = e;
                sse += (w ? w[i] : 1.0) * e * e;
            }
        } else {   /* baseline: plain double gemv, like nalgebra's y - X * beta */
            for (int64_t i = 0; i < n; ++i) {
                const double* xi = X + i * K;
                double yh = 0.0;
                for (int j = 0; j < K; ++j) yh += xi[j] * beta[j];
                const double e = y[i] - yh;
                if (resid) resid[i] = e;
                sse += (w ? w[i] : 1.0) * e * e;
            }
        }
        if (!precise) { /* ols.rs:124-137, discarded by the caller exactly as in the reference */
            volatile double sink = sse / ((double)n - (double)K) * chol_inverse_trace(G, K);
            (void)sink;
        }
    }
    free(G); free(r);
    return ORC_OK;
}

void orc_yun(const orc_spec* s, double* beta, int32_t idx_shift_from, double* base_coeff) {
    for (int32_t v = 0; v < s->n_norm; ++v) { /* normalization.rs:13-49 */
        base_coeff[v] = 0.0; /* HashMap miss -> unwrap_or(0.0) at builder.rs:652-654 */
        const int32_t a = s->norm_off[v], b = s->norm_off[v + 1];
        if (a == b) continue;                 /* :22-24 */
        double sum = 0.0;
        for (int32_t t = a; t < b; ++t) {     /* :26-29 */
            int32_t i = s->norm_idx[t];
            if (idx_shift_from >= 0 && i >= idx_shift_from) i += 1;
            sum += beta[i];
        }
        const int32_t m = s->norm_m[v];
        if (m == 0) continue;                 /* :35-37 */
        const double mu = sum / (double)m;    /* :38 */
        base_coeff[v] = -mu;                  /* :40 */
        beta[0] += mu;                        /* :44 */
        for (int32_t t = a; t < b; ++t) {     /* :46-48 */
            int32_t i = s->norm_idx[t];
            if (idx_shift_from >= 0 && i >= idx_shift_from) i += 1;
            beta[i] -= mu;
        }
    }
}

static double dotk(const double* a, const double* b, int K) {
    double s = 0.0;
    for (int i = 0; i < K; ++i) s += a[i] * b[i];
    return s;
}

void orc_two_fold(const double* xa, const double* xb, const double* ba, const double* bb,
                  const double* bs, int32_t K, double out[2]) { /* decomposition.rs:56-70 */
    double e = 0.0;
    for (int i = 0; i < K; ++i) e += (xa[i] - xb[i]) * bs[i];
    const double gap = dotk(xa, ba, K) - dotk(xb, bb, K);
    out[0] = e;
    out[1] = gap - e;
}

void orc_three_fold(const double* xa, const double* xb, const double* ba, const double* bb,
                    int32_t K, double out[3]) { /* decomposition.rs:73-89 */
    double en = 0.0, co = 0.0, in = 0.0;
    for (int i = 0; i < K; ++i) {
        const double dx = xa[i] - xb[i], db = ba[i] - bb[i];
        en += dx * bb[i];
        co += xb[i] * db;
        in += dx * db;
    }
    out[0] = en; out[1] = co; out[2] = in;
}

void orc_detailed(const double* xa, const double* xb, const double* ba, const double* bb,
                  const double* bs, int32_t K, double* expl, double* unexpl) { /* decomposition.rs:92-122 */
    for (int i = 0; i < K; ++i) {
        expl[i] = (xa[i] - xb[i]) * bs[i];
        unexpl[i] = xa[i] * (ba[i] - bs[i]) + xb[i] * (bs[i] - bb[i]);
    }
}

static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

void orc_bootstrap_stats(const double* est, int64_t n, double out[4]) { /* inference.rs:4-34 */
    if (n == 0) { out[0] = out[1] = out[2] = out[3] = NAN; return; }
    const double nf = (double)n;
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += est[i];
    const double mean = sum / nf;
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) { const double d = est[i] - mean; ss += d * d; }
    out[0] = sqrt(ss / (nf - 1.0));
    int64_t pos = 0, neg = 0;
    for (int64_t i = 0; i < n; ++i) { pos += est[i] >= 0.0; neg += est[i] <= 0.0; }
    const double pp = (double)pos / nf, pn = (double)neg / nf;
    out[1] = fmin(2.0 * fmin(pp, pn), 1.0);
    double* sorted = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(sorted, est, sizeof(double) * (size_t)n);
    qsort(sorted, (size_t)n, sizeof(double), cmp_double);
    int64_t lo = (int64_t)floor(0.025 * nf);
    int64_t hi = (int64_t)floor(0.975 * nf);
    if (hi > n - 1) hi = n - 1;
    out[2] = (lo < n) ? sorted[lo] : NAN;
    out[3] = sorted[hi];
    free(sorted);
}

void orc_rif(const double* y, int64_t n, double tau, double* rif_out) { /* math/rif.rs:14-88 */
    const double nf = (double)n;
    if (nf < 2.0) { memcpy(rif_out, y, sizeof(double) * (size_t)n); return; } /* :18-20 */
    double* s = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(s, y, sizeof(double) * (size_t)n);
    qsort(s, (size_t)n, sizeof(double), cmp_double);
    const double h = (nf - 1.0) * tau;                       /* :25 */
    const double hf = floor(h), hc = ceil(h), frac = h - hf;
    double q;
    if (hf == hc) q = s[(int64_t)hf];
    else { const double y0 = s[(int64_t)hf], y1 = s[(int64_t)hc]; q = y0 + frac * (y1 - y0); }
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += y[i];
    const double mean = sum / nf;                            /* :39 */
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) { const double d = y[i] - mean; ss += d * d; }
    const double sd = sqrt(ss / (nf - 1.0));                 /* :40-41 */
    int64_t i75 = (int64_t)ceil(0.75 * nf); i75 = i75 == 0 ? 0 : i75 - 1;   /* :43-44 */
    int64_t i25 = (int64_t)ceil(0.25 * nf); i25 = i25 == 0 ? 0 : i25 - 1;   /* :46-47 */
    if (i75 > n - 1) i75 = n - 1;
    if (i25 > n - 1) i25 = n - 1;
    const double iqr = s[i75] - s[i25];                      /* :49 */
    double spread = (iqr > 1e-8) ? fmin(sd, iqr / 1.34) : sd; /* :51-55 */
    if (spread < 1e-8) spread = 1.0;                         /* :57 */
    const double bw = 0.9 * spread * pow(nf, -0.2);          /* :59 */
    const double inv_sqrt_2pi = 1.0 / sqrt(2.0 * M_PI);
    double dens = 0.0;
    for (int64_t i = 0; i < n; ++i) {                        /* :65-72 */
        const double u = (q - y[i]) / bw;
        dens += inv_sqrt_2pi * exp(-0.5 * (u * u));
    }
    dens /= (nf * bw);
    if (dens < 1e-8) dens = 1e-8;                            /* :75 */
    for (int64_t i = 0; i < n; ++i) {                        /* :79-85 */
        const double ind = (y[i] <= q) ? 1.0 : 0.0;
        rif_out[i] = q + (tau - ind) / dens;
    }
    free(s);
}

/* estimation.rs:56-68: weighted mean sum(w x)/sum(w) or plain column mean */
static void col_means(const double* X, const double* w, int64_t n, int K, int precise, double* mean) {
    if (precise) {
        ld* acc = (ld*)calloc((size_t)K, sizeof(ld));
        ld tw = 0;
        for (int64_t i = 0; i < n; ++i) {
            const ld wi = w ? (ld)w[i] : 1.0L;
            tw += wi;
            for (int j = 0; j < K; ++j) acc[j] += wi * (ld)X[i * K + j];
        }
        for (int j = 0; j < K; ++j) mean[j] = (double)(acc[j] / (w ? tw : (ld)n));
        free(acc);
    } else {
        double* acc = (double*)calloc((size_t)K, sizeof(double));
        double tw = 0;
        for (int64_t i = 0; i < n; ++i) {
            const double wi = w ? w[i] : 1.0;
            tw += wi;
            for (int j = 0; j < K; ++j) acc[j] += wi * X[i * K + j];
        }
        for (int j = 0; j < K; ++j) mean[j] = acc[j] / (w ? tw : (double)n);
        free(acc);
    }
}

static double wmean(const double* y, const double* w, int64_t n) { /* builder.rs:676-684 */
    ld s = 0, tw = 0;
    for (int64_t i = 0; i < n; ++i) {
        const ld wi = w ? (ld)w[i] : 1.0L;
        s += wi * (ld)y[i];
        tw += wi;
    }
    return (double)(s / (w ? tw : (ld)n));
}

int orc_single_pass(const orc_spec* s,
                    const double* Xa, const double* ya, const double* wa, int64_t na,
                    const double* Xb, const double* yb, const double* wb, int64_t nb,
                    int precise, orc_pass_out* out) {
    const int K = s->K;
    if (na == 0 || nb == 0) return ORC_ERR_INVALID_GROUP; /* builder.rs:431-435 */
    int rc;
    /* estimation.rs:53-54 */
    if ((rc = orc_ols(ya, Xa, wa, na, K, precise, out->beta_a, out->resid_a)) != ORC_OK) return rc;
    if ((rc = orc_ols(yb, Xb, wb, nb, K, precise, out->beta_b, out->resid_b)) != ORC_OK) return rc;
    col_means(Xa, wa, na, K, precise, out->xa_mean); /* estimation.rs:70-71 */
    col_means(Xb, wb, nb, K, precise, out->xb_mean);

    double* base_a = (double*)calloc((size_t)(s->n_norm + 1), sizeof(double));
    double* base_b = (double*)calloc((size_t)(s->n_norm + 1), sizeof(double));
    double* base_s = (double*)calloc((size_t)(s->n_norm + 1), sizeof(double));
    if (s->n_norm > 0) { /* estimation.rs:76-91 */
        orc_yun(s, out->beta_a, -1, base_a);
        orc_yun(s, out->beta_b, -1, base_b);
    }

    /* beta* : builder.rs:538-621 */
    switch (s->ref_kind) {
    case ORC_REF_GROUP_A:
        memcpy(out->beta_star, out->beta_a, sizeof(double) * (size_t)K);
        memcpy(base_s, base_a, sizeof(double) * (size_t)s->n_norm);
        break;
    case ORC_REF_GROUP_B:
        memcpy(out->beta_star, out->beta_b, sizeof(double) * (size_t)K);
        memcpy(base_s, base_b, sizeof(double) * (size_t)s->n_norm);
        break;
    case ORC_REF_POOLED: {
        /* builder.rs:548-589: vstack(A,B), indicator (1 for A) appended to the *predictors*,
         * hence placed after the continuous ones and before the dummies (:322-327) */
        const int Kp = K + 1, ind = 1 + s->n_cont;
        const int64_t n = na + nb;
        double* Xp = (double*)malloc(sizeof(double) * (size_t)n * Kp);
        double* yp = (double*)malloc(sizeof(double) * (size_t)n);
        double* wp = (wa && wb) ? (double*)malloc(sizeof(double) * (size_t)n) : NULL;
        for (int64_t i = 0; i < n; ++i) {
            const int isa = i < na;
            const double* src = isa ? Xa + i * K : Xb + (i - na) * K;
            double* dst = Xp + i * Kp;
            for (int c = 0; c < ind; ++c) dst[c] = src[c];
            dst[ind] = isa ? 1.0 : 0.0;
            for (int c = ind; c < K; ++c) dst[c + 1] = src[c];
            yp[i] = isa ? ya[i] : yb[i - na];
            if (wp) wp[i] = isa ? wa[i] : wb[i - na];
        }
        double* bp = (double*)malloc(sizeof(double) * (size_t)Kp);
        rc = orc_ols(yp, Xp, wp, n, Kp, precise, bp, NULL); /* :566 */
        free(Xp); free(yp); free(wp);
        if (rc != ORC_OK) { free(bp); free(base_a); free(base_b); free(base_s); return rc; }
        if (s->n_norm > 0) orc_yun(s, bp, ind, base_s);     /* :568-579 with pooled names */
        for (int c = 0; c < ind; ++c) out->beta_star[c] = bp[c];       /* :580-589 remove_row */
        for (int c = ind; c < K; ++c) out->beta_star[c] = bp[c + 1];
        free(bp);
        break;
    }
    default: { /* Weighted | Cotton : builder.rs:591-620 */
        double n_a = (double)na, n_b = (double)nb;
        if (wa) { n_a = 0.0; for (int64_t i = 0; i < na; ++i) n_a += wa[i]; }
        if (wb) { n_b = 0.0; for (int64_t i = 0; i < nb; ++i) n_b += wb[i]; }
        const double total = n_a + n_b;
        if (total == 0.0) { free(base_a); free(base_b); free(base_s); return ORC_ERR_INVALID_GROUP; }
        const double wA = n_a / total, wB = 1.0 - wA;
        for (int32_t v = 0; v < s->n_norm; ++v) base_s[v] = base_a[v] * wA + base_b[v] * wB;
        for (int c = 0; c < K; ++c) out->beta_star[c] = out->beta_a[c] * wA + out->beta_b[c] * wB;
        break;
    }
    }

    /* builder.rs:623-632 */
    orc_three_fold(out->xa_mean, out->xb_mean, out->beta_a, out->beta_b, K, out->three_fold);
    orc_two_fold(out->xa_mean, out->xb_mean, out->beta_a, out->beta_b, out->beta_star, K, out->two_fold);
    orc_detailed(out->xa_mean, out->xb_mean, out->beta_a, out->beta_b, out->beta_star, K,
                 out->det_expl, out->det_unexpl);

    /* Yun base-category rows: builder.rs:634-674.  three_fold is NOT corrected. */
    int row = K;
    for (int32_t v = 0; v < s->n_norm; ++v) {
        if (!s->norm_has_base[v]) continue;
        double sa = 0.0, sb = 0.0;
        for (int32_t t = s->norm_off[v]; t < s->norm_off[v + 1]; ++t) {
            sa += out->xa_mean[s->norm_idx[t]];
            sb += out->xb_mean[s->norm_idx[t]];
        }
        const double xab = 1.0 - sa, xbb = 1.0 - sb;
        const double un = xab * (base_a[v] - base_s[v]) + xbb * (base_s[v] - base_b[v]);
        const double ex = (xab - xbb) * base_s[v];
        out->det_unexpl[row] = un;
        out->det_expl[row] = ex;
        out->two_fold[0] += ex;
        out->two_fold[1] += un;
        ++row;
    }
    out->total_gap = wmean(ya, wa, na) - wmean(yb, wb, nb);
    free(base_a); free(base_b); free(base_s);
    return ORC_OK;
}

void orc_pass_to_stats(const orc_spec* s, const orc_pass_out* p, double* stats) {
    const int D = s->K + orc_n_base(s);
    stats[0] = p->two_fold[0]; stats[1] = p->two_fold[1];
    stats[2] = p->three_fold[0]; stats[3] = p->three_fold[1]; stats[4] = p->three_fold[2];
    memcpy(stats + 5, p->det_expl, sizeof(double) * (size_t)D);
    memcpy(stats + 5 + D, p->det_unexpl, sizeof(double) * (size_t)D);
}

/* ---- the oracle's own resampling stream: xoshiro256** seeded by splitmix64(seed, rep, group) ---- */
static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

void orc_fill_indices(uint64_t seed, int64_t rep, int32_t group, int64_t n, uint32_t* idx) {
    uint64_t sm = seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(rep + 1)) ^ (0x8CB92BA72F3D8DD7ULL * (uint64_t)(group + 1));
    uint64_t st[4];
    for (int i = 0; i < 4; ++i) st[i] = splitmix64(&sm);
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t r = rotl64(st[1] * 5, 7) * 9;
        const uint64_t t = st[1] << 17;
        st[2] ^= st[0]; st[3] ^= st[1]; st[1] ^= st[2]; st[0] ^= st[3]; st[2] ^= t; st[3] = rotl64(st[3], 45);
        /* Uniform(0, n): multiply-shift on the top 32 bits, rejection-free (bias < 2^-32 * n) */
        idx[i] = (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
    }
}

void orc_reduce(const double* rep_stats, const int32_t* rep_status, int64_t reps, int32_t S,
                const double* point_stats, int64_t* n_ok_out,
                double* se, double* p, double* ci_lo, double* ci_hi, double* t) {
    int64_t n_ok = 0;
    for (int64_t b = 0; b < reps; ++b) n_ok += (rep_status[b] == ORC_OK);
    double* col = (double*)malloc(sizeof(double) * (size_t)(n_ok > 0 ? n_ok : 1));
    for (int32_t j = 0; j < S; ++j) {
        int64_t m = 0;
        for (int64_t b = 0; b < reps; ++b) /* filter_map keeps replicate order: builder.rs:816-839 */
            if (rep_status[b] == ORC_OK) col[m++] = rep_stats[b * S + j];
        double o[4];
        orc_bootstrap_stats(col, n_ok, o);
        se[j] = o[0]; p[j] = o[1]; ci_lo[j] = o[2]; ci_hi[j] = o[3];
        t[j] = (fabs(o[0]) > 1e-9) ? point_stats[j] / o[0] : 0.0; /* builder.rs:851-855 */
    }
    free(col);
    *n_ok_out = n_ok;
}

int orc_run(const orc_spec* s,
            const double* Xa, const double* ya, const double* wa, int64_t na,
            const double* Xb, const double* yb, const double* wb, int64_t nb,
            int64_t reps, const uint32_t* idx_a, const uint32_t* idx_b, uint64_t seed,
            int nthreads, int precise, orc_run_out* out) {
    const int K = s->K, S = orc_n_stats(s), D = K + orc_n_base(s);
    /* point estimates: builder.rs:810-811; a failure here is a hard error.  The pass runs as work item -1 of the
     * replicate loop below (the point pass and the replicates are independent), so that a timed run of T-1 replicates
     * on T threads costs one pass time, not two. */
    int rc = ORC_OK;
    if (nthreads < 1) nthreads = 1;

#pragma omp parallel num_threads(nthreads)
    {
        double* Xsa = (double*)malloc(sizeof(double) * (size_t)na * K);
        double* Xsb = (double*)malloc(sizeof(double) * (size_t)nb * K);
        double* ysa = (double*)malloc(sizeof(double) * (size_t)na);
        double* ysb = (double*)malloc(sizeof(double) * (size_t)nb);
        double* wsa = wa ? (double*)malloc(sizeof(double) * (size_t)na) : NULL;
        double* wsb = wb ? (double*)malloc(sizeof(double) * (size_t)nb) : NULL;
        uint32_t* ia = idx_a ? NULL : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)na);
        uint32_t* ib = idx_b ? NULL : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nb);
        double* buf = (double*)malloc(sizeof(double) * (size_t)(2 * D + 5 * K));
        orc_pass_out po;
        memset(&po, 0, sizeof(po));
        po.det_expl = buf; po.det_unexpl = buf + D;
        po.xa_mean = buf + 2 * D; po.xb_mean = po.xa_mean + K; po.beta_star = po.xb_mean + K;
        po.beta_a = po.beta_star + K; po.beta_b = po.beta_a + K;
        double* st = (double*)malloc(sizeof(double) * (size_t)S);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = -1; b < reps; ++b) {
            if (b < 0) { rc = orc_single_pass(s, Xa, ya, wa, na, Xb, yb, wb, nb, precise, &out->point); continue; }
            /* builder.rs:822-829: sample_n_literal(height, with_replacement) per group, vstack */
            const uint32_t* ja = idx_a ? idx_a + b * na : ia;
            const uint32_t* jb = idx_b ? idx_b + b * nb : ib;
            if (!idx_a) orc_fill_indices(seed, b, 0, na, ia);
            if (!idx_b) orc_fill_indices(seed, b, 1, nb, ib);
            for (int64_t i = 0; i < na; ++i) {
                memcpy(Xsa + i * K, Xa + (int64_t)ja[i] * K, sizeof(double) * (size_t)K);
                ysa[i] = ya[ja[i]];
                if (wsa) wsa[i] = wa[ja[i]];
            }
            for (int64_t i = 0; i < nb; ++i) {
                memcpy(Xsb + i * K, Xb + (int64_t)jb[i] * K, sizeof(double) * (size_t)K);
                ysb[i] = yb[jb[i]];
                if (wsb) wsb[i] = wb[jb[i]];
            }
            orc_reset_min_pivot();
            const int r = orc_single_pass(s, Xsa, ysa, wsa, na, Xsb, ysb, wsb, nb, precise, &po);
            if (out->rep_status) out->rep_status[b] = r;
            if (out->rep_min_pivot) out->rep_min_pivot[b] = orc_last_min_pivot();
            if (r == ORC_OK) {
                orc_pass_to_stats(s, &po, st);
                if (out->rep_stats) memcpy(out->rep_stats + b * S, st, sizeof(double) * (size_t)S);
                if (out->rep_beta_a) memcpy(out->rep_beta_a + b * K, po.beta_a, sizeof(double) * (size_t)K);
                if (out->rep_beta_b) memcpy(out->rep_beta_b + b * K, po.beta_b, sizeof(double) * (size_t)K);
            } else { /* .ok()? -> dropped: builder.rs:831-837 */
                if (out->rep_stats) for (int j = 0; j < S; ++j) out->rep_stats[b * S + j] = NAN;
                if (out->rep_beta_a) for (int j = 0; j < K; ++j) out->rep_beta_a[b * K + j] = NAN;
                if (out->rep_beta_b) for (int j = 0; j < K; ++j) out->rep_beta_b[b * K + j] = NAN;
            }
        }
        free(Xsa); free(Xsb); free(ysa); free(ysb); free(wsa); free(wsb); free(ia); free(ib);
        free(buf); free(st);
    }

    if (rc != ORC_OK) return rc;
    if (out->rep_stats && out->rep_status && out->se) {
        double* pst = (double*)malloc(sizeof(double) * (size_t)S);
        orc_pass_to_stats(s, &out->point, pst);
        orc_reduce(out->rep_stats, out->rep_status, reps, S, pst, &out->n_ok,
                   out->se, out->p, out->ci_lo, out->ci_hi, out->t);
        free(pst);
    }
    return ORC_OK;
}
