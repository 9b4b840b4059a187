/* oracle/ob_oracle_mm.c -- CPU restatement of the Machado-Mata quantile decomposition.
 *
 * TEST INFRASTRUCTURE ONLY (see ob_oracle.h).  Citations are file:line under oaxaca_blinder/src/ of the reference.
 *
 * The reference solves every quantile regression as the LP
 *     min  tau 1'u + (1 - tau) 1'v   s.t.  X beta + u - v = y,  u, v >= 0          (math/quantile_regression.rs:22-135)
 * with the interior-point conic solver `clarabel` (Cargo dependency, not vendored, not buildable here: no Rust
 * toolchain), at its default tolerances (~1e-8): its coefficients are the LP's solution only up to that tolerance.
 * What IS defined independently of any solver is the LP's optimal vertex (unique for data in general position: the
 * hyperplane through K observations).  This oracle computes that vertex: a Frisch-Newton primal-dual interior-point
 * method (Portnoy & Koenker 1997, the algorithm behind quantreg's rq.fit.fnb) on the bounded dual
 *     max y'a   s.t.  X'a = (1 - tau) X'1,  0 <= a <= 1,
 * run to a relative duality gap of 1e-12, followed by a polish: the observations with a (numerically) zero residual
 * are identified, beta is refined to the exact solution of X_h beta = y_h over them (long double), and the result is
 * verified (zero residuals on h, unchanged residual signs elsewhere).  Pinned against an independent solver
 * (scipy.optimize.linprog / HiGHS dual simplex: tests/golden/make_mm_golden.py -> mm_fixture.json) and against the
 * two known-answer tests the reference holds (quantile_regression.rs:137-170).  The random quantiles and the simulated
 * rows are unseeded in the reference (thread_rng, quantile_decomposition.rs:221-225, :251-255), so like the bootstrap
 * resamples they are harness-provided streams here: "parity unpinned" in the reference itself.
 */
#include "ob_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;

/* in-place lower Cholesky of a K x K row-major matrix; 0 = ok */
static int mm_chol(double* G, int K) {
    for (int j = 0; j < K; ++j) {
        const double d0 = G[j * K + j];
        double d = d0;
        for (int k = 0; k < j; ++k) d -= G[j * K + k] * G[j * K + k];
        if (!(d > 1e-14 * d0) || !isfinite(d)) return 1;      /* rank-deficient up to rounding (n < K, collinear columns) */
        d = sqrt(d);
        G[j * K + j] = d;
        for (int i = j + 1; i < K; ++i) {
            double s = G[i * K + j];
            for (int k = 0; k < j; ++k) s -= G[i * K + k] * G[j * K + k];
            G[i * K + j] = s / d;
        }
    }
    return 0;
}
static void mm_chol_solve(const double* L, int K, double* b) {
    for (int i = 0; i < K; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * K + k] * b[k];
        b[i] = s / L[i * K + i];
    }
    for (int i = K - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < K; ++k) s -= L[k * K + i] * b[k];
        b[i] = s / L[i * K + i];
    }
}
static int mm_chol_ld(ld* G, int K) {
    for (int j = 0; j < K; ++j) {
        const ld d0 = G[j * K + j];
        ld d = d0;
        for (int k = 0; k < j; ++k) d -= G[j * K + k] * G[j * K + k];
        if (!(d > 1e-12L * d0)) return 1;       /* numerically rank-deficient */
        d = sqrtl(d);
        G[j * K + j] = d;
        for (int i = j + 1; i < K; ++i) {
            ld s = G[i * K + j];
            for (int k = 0; k < j; ++k) s -= G[i * K + k] * G[j * K + k];
            G[i * K + j] = s / d;
        }
    }
    return 0;
}
static void mm_chol_solve_ld(const ld* L, int K, ld* b) {
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

/* G = sum_i q_i x_i x_i' (lower part filled, symmetric), g = sum_i q_i v_i x_i */
static void mm_gram(const double* X, const double* q, const double* v, int64_t n, int K, double* G, double* g) {
    memset(G, 0, sizeof(double) * (size_t)K * K);
    memset(g, 0, sizeof(double) * (size_t)K);
    for (int64_t i = 0; i < n; ++i) {
        const double qi = q[i];
        if (qi == 0.0) continue;
        const double* x = X + i * K;
        const double qv = qi * v[i];
        for (int j = 0; j < K; ++j) {
            const double a = qi * x[j];
            for (int l = 0; l <= j; ++l) G[j * K + l] += a * x[l];
            g[j] += qv * x[j];
        }
    }
    for (int j = 0; j < K; ++j)
        for (int l = j + 1; l < K; ++l) G[j * K + l] = G[l * K + j];
}

/* g = sum_i q_i v_i x_i */
static void mm_xtqv(const double* X, const double* q, const double* v, int64_t n, int K, double* g) {
    memset(g, 0, sizeof(double) * (size_t)K);
    for (int64_t i = 0; i < n; ++i) {
        const double qv = q[i] * v[i];
        if (q[i] == 0.0) continue;
        for (int j = 0; j < K; ++j) g[j] += qv * X[i * K + j];
    }
}

static double mm_dot(const double* a, const double* b, int K) {
    double s = 0.0;
    for (int j = 0; j < K; ++j) s += a[j] * b[j];
    return s;
}

/* info: [0] IPM iterations, [1] status (0 = vertex verified, 1 = interior-point solution only, 2 = failed),
 *       [2] zero-residual observations used by the polish */
enum { QR_VERTEX = 0, QR_APPROX = 1, QR_FAILED = 2 };

/* math/quantile_regression.rs:22-135: coefficients of the tau-th regression quantile of y on X [n x K] (row-major).
 * c: optional observation multiplicities (>= 0; NULL = 1): min sum_i c_i rho_tau(y_i - x_i'beta), the same LP as the
 * regression on a frame in which row i occurs c_i times. */
int orc_qr(const double* X, const double* y, const double* c, int64_t n, int32_t K, double tau, double* beta, int32_t* info) {
    int32_t dummy[3];
    if (!info) info = dummy;
    info[0] = 0; info[1] = QR_FAILED; info[2] = 0;
    if (n < 1 || K < 1 || !(tau >= 0.0 && tau <= 1.0)) return 1;      /* quantile_regression.rs:30-32 */
    tau = fmin(fmax(tau, 1e-6), 1.0 - 1e-6);        /* the ends of [0, 1] keep an interior starting point */
    double* buf = (double*)malloc(sizeof(double) * (size_t)n * 9);
    double *u = buf, *x = u + n, *s = x + n, *z = s + n, *w = z + n, *q = w + n, *r = q + n, *dxa = r + n, *xi = dxa + n;
    double* G = (double*)malloc(sizeof(double) * (size_t)K * (K + 4));
    double *g = G + (size_t)K * K, *yd = g + K, *dyd = yd + K, *g2 = dyd + K;
    int rc = 1;
    double yscale = 0.0, scale = 0.0, usum = 0.0;
    int64_t nact = 0;
    for (int64_t i = 0; i < n; ++i) {
        u[i] = c ? c[i] : 1.0;
        if (u[i] > 0.0) { ++nact; if (fabs(y[i]) > yscale) yscale = fabs(y[i]); scale += u[i] * fabs(y[i]); usum += u[i]; }
    }
    if (nact < 1) goto done;
    if (yscale == 0.0) { yscale = 1.0; }
    if (scale == 0.0) scale = yscale;
    /* starting point: x = (1 - tau) u (primal feasible), yd = least-squares fit of the cost -y, z - w = -y - X yd */
    for (int64_t i = 0; i < n; ++i) r[i] = -y[i];
    mm_gram(X, u, r, n, K, G, g);
    if (mm_chol(G, K)) goto done;
    memcpy(yd, g, sizeof(double) * (size_t)K);
    mm_chol_solve(G, K, yd);
    double mean_abs = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        r[i] = -y[i] - mm_dot(X + i * K, yd, K);
        if (u[i] > 0.0) mean_abs += u[i] * fabs(r[i]);
    }
    mean_abs /= usum;
    const double delta = 0.01 * mean_abs + 1e-10 * yscale;
    double gap = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (!(u[i] > 0.0)) { x[i] = s[i] = z[i] = w[i] = q[i] = 0.0; continue; }
        x[i] = (1.0 - tau) * u[i]; s[i] = u[i] - x[i];
        z[i] = fmax(r[i], 0.0) + delta; w[i] = fmax(-r[i], 0.0) + delta;
        gap += z[i] * x[i] + w[i] * s[i];
    }
    const double tol = 1e-12 * scale;
    int it = 0;
    for (; it < 100 && gap > tol; ++it) {
        for (int64_t i = 0; i < n; ++i) {
            if (!(u[i] > 0.0)) continue;
            q[i] = 1.0 / (z[i] / x[i] + w[i] / s[i]);
            r[i] = z[i] - w[i];
        }
        mm_gram(X, q, r, n, K, G, g);
        if (mm_chol(G, K)) break;
        memcpy(dyd, g, sizeof(double) * (size_t)K);
        mm_chol_solve(G, K, dyd);
        /* affine step, ratio test, and the sums that give the gap after the step for any pair of step lengths */
        double ap = 1e300, ad = 1e300, S1 = 0.0, S3 = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            if (!(u[i] > 0.0)) continue;
            const double dx = q[i] * (mm_dot(X + i * K, dyd, K) - r[i]);
            const double dz = -z[i] * (1.0 + dx / x[i]), dw = -w[i] * (1.0 - dx / s[i]);
            dxa[i] = dx;
            if (dx < 0.0) ap = fmin(ap, -x[i] / dx);
            if (dx > 0.0) ap = fmin(ap, s[i] / dx);
            if (dz < 0.0) ad = fmin(ad, -z[i] / dz);
            if (dw < 0.0) ad = fmin(ad, -w[i] / dw);
            S1 += dx * r[i]; S3 += dx * (dz - dw);
        }
        ap = fmin(0.99995 * ap, 1.0); ad = fmin(0.99995 * ad, 1.0);
        int corrector = 0;
        double mu = 0.0;
        if (fmin(ap, ad) < 1.0) {
            corrector = 1;
            const double gaff = gap + ap * S1 + ad * (-gap - S1) + ap * ad * S3;
            const double ratio = gaff / gap;
            mu = gap * ratio * ratio * ratio / (2.0 * (double)nact);
            if (!(mu >= 0.0)) mu = 0.0;
            for (int64_t i = 0; i < n; ++i) {
                if (!(u[i] > 0.0)) { xi[i] = 0.0; continue; }
                const double dx = dxa[i];
                const double dz = -z[i] * (1.0 + dx / x[i]), dw = -w[i] * (1.0 - dx / s[i]);
                /* xi = r + mu (1/s - 1/x) + dx dz / x - ds dw / s, ds = -dx */
                xi[i] = r[i] + mu * (1.0 / s[i] - 1.0 / x[i]) + dx * dz / x[i] + dx * dw / s[i];
            }
            mm_xtqv(X, q, xi, n, K, g2);
            memcpy(dyd, g2, sizeof(double) * (size_t)K);
            mm_chol_solve(G, K, dyd);
            ap = 1e300; ad = 1e300;
            for (int64_t i = 0; i < n; ++i) {
                if (!(u[i] > 0.0)) continue;
                const double dxaff = dxa[i];
                const double dzaff = -z[i] * (1.0 + dxaff / x[i]), dwaff = -w[i] * (1.0 - dxaff / s[i]);
                const double dx = q[i] * (mm_dot(X + i * K, dyd, K) - xi[i]);
                const double dz = (mu - dxaff * dzaff) / x[i] - z[i] - z[i] / x[i] * dx;
                const double dw = (mu + dxaff * dwaff) / s[i] - w[i] + w[i] / s[i] * dx;
                if (dx < 0.0) ap = fmin(ap, -x[i] / dx);
                if (dx > 0.0) ap = fmin(ap, s[i] / dx);
                if (dz < 0.0) ad = fmin(ad, -z[i] / dz);
                if (dw < 0.0) ad = fmin(ad, -w[i] / dw);
                xi[i] = dx;             /* final dx */
                r[i] = dz; q[i] = dw;   /* final dz, dw (r and q are rebuilt next iteration) */
            }
            ap = fmin(0.99995 * ap, 1.0); ad = fmin(0.99995 * ad, 1.0);
        }
        /* one step length for the primal and the dual iterate: with separate ones the iterates of the extreme quantiles
         * (tau near 0.01 / 0.99, where all but a few percent of the dual variables end at one bound) lose centrality and
         * the iteration count grows from ~20 to 100 and more */
        ap = ad = fmin(ap, ad);
        gap = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            if (!(u[i] > 0.0)) continue;
            double dx, dz, dw;
            if (corrector) { dx = xi[i]; dz = r[i]; dw = q[i]; }
            else { dx = dxa[i]; dz = -z[i] * (1.0 + dx / x[i]); dw = -w[i] * (1.0 - dx / s[i]); }
            x[i] += ap * dx; s[i] = u[i] - x[i];
            z[i] += ad * dz; w[i] += ad * dw;
            gap += z[i] * x[i] + w[i] * s[i];
        }
        for (int j = 0; j < K; ++j) yd[j] += ad * dyd[j];
        if (!isfinite(gap)) break;
        if (ap < 1e-12 && ad < 1e-12) { ++it; break; }
    }
    info[0] = it;
    if (!isfinite(gap) || gap > 1e-7 * scale) goto done;       /* "Solver failed" (quantile_regression.rs:127-132) */
    for (int j = 0; j < K; ++j) beta[j] = -yd[j];
    info[1] = QR_APPROX;
    rc = 0;
    {   /* polish to the LP's vertex */
        ld* res = (ld*)malloc(sizeof(ld) * (size_t)n);
        int64_t cnt[16] = {0};
        for (int64_t i = 0; i < n; ++i) {
            if (!(u[i] > 0.0)) continue;
            ld a = y[i];
            for (int j = 0; j < K; ++j) a -= (ld)X[i * K + j] * (ld)beta[j];
            res[i] = a;
            double t = yscale * 1e-3;
            for (int j = 3; j <= 14; ++j, t *= 0.1)
                if (fabsl(a) < t) ++cnt[j];
        }
        int jstar = -1;
        for (int j = 14; j >= 3; --j) if (cnt[j] >= K) { jstar = j; break; }
        /* candidates = the rows within a decade of the smallest threshold that K rows pass; if those do not span R^K
         * (duplicated rows of a resample, a basic row the interior-point iterate has not reached as closely) or the
         * refined point fails the verification, the next decade is tried */
        ld* Gh = (ld*)calloc((size_t)K * K + 2 * (size_t)K, sizeof(ld));
        ld *gh = Gh + (size_t)K * K, *bl = gh + K;
        unsigned char* cand = (unsigned char*)calloc((size_t)n, 1);
        int64_t m_prev = -1;
        for (int jt = jstar; jt >= 4 && info[1] != QR_VERTEX; --jt) {
            const double thr = yscale * pow(10.0, -(double)(jt - 1));
            int64_t m = 0;
            for (int j = 0; j < K * K; ++j) Gh[j] = 0;
            for (int64_t i = 0; i < n; ++i) {
                cand[i] = 0;
                if (!(u[i] > 0.0) || !(fabsl(res[i]) < thr)) continue;
                cand[i] = 1; ++m;
                for (int j = 0; j < K; ++j)
                    for (int l = 0; l <= j; ++l) Gh[j * K + l] += (ld)X[i * K + j] * (ld)X[i * K + l];
            }
            if (m == m_prev) continue;
            m_prev = m;
            int ok = mm_chol_ld(Gh, K) == 0;
            for (int j = 0; j < K; ++j) bl[j] = beta[j];
            for (int round = 0; ok && round < 3; ++round) {
                for (int j = 0; j < K; ++j) gh[j] = 0;
                for (int64_t i = 0; i < n; ++i) {
                    if (!cand[i]) continue;
                    ld a = y[i];
                    for (int j = 0; j < K; ++j) a -= (ld)X[i * K + j] * bl[j];
                    for (int j = 0; j < K; ++j) gh[j] += (ld)X[i * K + j] * a;
                }
                mm_chol_solve_ld(Gh, K, gh);
                for (int j = 0; j < K; ++j) bl[j] += gh[j];
            }
            for (int64_t i = 0; ok && i < n; ++i) {
                if (!(u[i] > 0.0)) continue;
                ld a = y[i];
                for (int j = 0; j < K; ++j) a -= (ld)X[i * K + j] * bl[j];
                if (cand[i]) { if (fabsl(a) > 1e-11L * yscale) ok = 0; }
                else if ((a > 0) != (res[i] > 0)) ok = 0;
            }
            if (ok) {
                for (int j = 0; j < K; ++j) beta[j] = (double)bl[j];
                info[1] = QR_VERTEX; info[2] = (int32_t)m;
            }
        }
        free(Gh); free(cand);
        free(res);
    }
done:
    free(buf); free(G);
    return rc;
}

static int mm_cmp(const void* a, const void* b) {      /* partial_cmp().unwrap_or(Equal), quantile_decomposition.rs:168 */
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}
/* quantile_decomposition.rs:164-171 */
static double mm_empirical_quantile(double* data, int64_t len, double quantile) {
    if (len == 0) return 0.0;
    qsort(data, (size_t)len, sizeof(double), mm_cmp);
    int64_t index = (int64_t)((double)len * quantile);
    if (index > len - 1) index = len - 1;
    return data[index];
}

/* run_single_pass, quantile_decomposition.rs:173-279, on the split dense designs (X incl. the intercept column).
 * taus [sims]: the pass's random quantiles (:221-225); draw_a / draw_b [sims]: the simulated row of each group (:251-255).
 * stats [3 nq]: (gap, characteristics, coefficients) per target quantile (:268-276).  Optional outputs: betas_* [sims x K]
 * (rows of failed regressions NaN), qr_status_* [sims] (QR_*), nsucc. */
int orc_mm_pass(int32_t K, const double* Xa, const double* ya, int64_t na, const double* Xb, const double* yb, int64_t nb,
                int32_t sims, const double* taus, const uint32_t* draw_a, const uint32_t* draw_b,
                int32_t nq, const double* quantiles, double* stats,
                double* betas_a, double* betas_b, int32_t* qr_status_a, int32_t* qr_status_b, int32_t* nsucc_out) {
    if (na < 2 || nb < 2) return ORC_ERR_INVALID_GROUP;                  /* :210-214 */
    double* ba = (double*)malloc(sizeof(double) * (size_t)sims * K * 2);
    double* bb = ba + (size_t)sims * K;
    int32_t* oka = (int32_t*)malloc(sizeof(int32_t) * (size_t)sims * 2);
    int32_t* okb = oka + sims;
    for (int32_t s = 0; s < sims; ++s) {                                 /* :227-236 (par_iter; order preserved) */
        int32_t info[3];
        orc_qr(Xa, ya, NULL, na, K, taus[s], ba + (size_t)s * K, info); oka[s] = info[1];
        orc_qr(Xb, yb, NULL, nb, K, taus[s], bb + (size_t)s * K, info); okb[s] = info[1];
    }
    for (int32_t s = 0; s < sims; ++s) {
        if (qr_status_a) qr_status_a[s] = oka[s];
        if (qr_status_b) qr_status_b[s] = okb[s];
        for (int j = 0; j < K; ++j) {
            if (betas_a) betas_a[(size_t)s * K + j] = oka[s] == QR_FAILED ? NAN : ba[(size_t)s * K + j];
            if (betas_b) betas_b[(size_t)s * K + j] = okb[s] == QR_FAILED ? NAN : bb[(size_t)s * K + j];
        }
    }
    /* filter_map(.ok()): the successful fits of each group, in order */
    int32_t la = 0, lb = 0;
    for (int32_t s = 0; s < sims; ++s) {
        if (oka[s] != QR_FAILED) { if (la != s) memcpy(ba + (size_t)la * K, ba + (size_t)s * K, sizeof(double) * (size_t)K); ++la; }
        if (okb[s] != QR_FAILED) { if (lb != s) memcpy(bb + (size_t)lb * K, bb + (size_t)s * K, sizeof(double) * (size_t)K); ++lb; }
    }
    int rc = ORC_OK;
    if (la < sims / 2 || lb < sims / 2) rc = ORC_ERR_NALGEBRA;           /* :238-242 */
    const int32_t ns = la < lb ? la : lb;                                /* :244 */
    if (nsucc_out) *nsucc_out = ns;
    if (rc == ORC_OK) {
        double* yv = (double*)malloc(sizeof(double) * (size_t)(ns > 0 ? ns : 1) * 3);
        double *yaa = yv, *ybb = yv + ns, *yab = yv + 2 * (size_t)ns;
        for (int32_t i = 0; i < ns; ++i) {                               /* :252-264 */
            const double* xa = Xa + (size_t)draw_a[i] * K;
            const double* xb = Xb + (size_t)draw_b[i] * K;
            yaa[i] = mm_dot(xa, ba + (size_t)i * K, K);
            ybb[i] = mm_dot(xb, bb + (size_t)i * K, K);
            yab[i] = mm_dot(xa, bb + (size_t)i * K, K);
        }
        for (int32_t k = 0; k < nq; ++k) {                               /* :266-277 */
            const double qaa = mm_empirical_quantile(yaa, ns, quantiles[k]);
            const double qbb = mm_empirical_quantile(ybb, ns, quantiles[k]);
            const double qab = mm_empirical_quantile(yab, ns, quantiles[k]);
            stats[3 * k + 0] = qaa - qbb;
            stats[3 * k + 1] = qab - qbb;
            stats[3 * k + 2] = qaa - qab;
        }
        free(yv);
    }
    free(ba); free(oka);
    return rc;
}

/* QuantileDecompositionBuilder::run, quantile_decomposition.rs:281-421, from the split designs on.
 * Streams (all unseeded in the reference): idx_a [reps x na], idx_b [reps x nb] the bootstrap resamples
 * (sample_n_literal, :343-348); taus [(reps + 1) x sims] and draw_a / draw_b [(reps + 1) x sims] the random quantiles and
 * simulated rows of every pass, pass 0 = the point estimates, draws = positions in the pass's (resampled) group frame.
 * Outputs: point_stats [3 nq]; rep_stats [reps x 3 nq] (NaN rows for failed passes), rep_status [reps]; reduction
 * (bootstrap_stats, :358-363; t = point / se when |se| > 1e-9, :371-375) se, p, ci_lo, ci_hi, t [3 nq]. */
int orc_mm_run(int32_t K, const double* Xa, const double* ya, int64_t na, const double* Xb, const double* yb, int64_t nb,
               int32_t sims, int32_t nq, const double* quantiles, int64_t reps,
               const uint32_t* idx_a, const uint32_t* idx_b, const double* taus, const uint32_t* draw_a, const uint32_t* draw_b,
               int nthreads, double* point_stats, double* point_betas_a, double* point_betas_b,
               double* rep_stats, int32_t* rep_status, int64_t* n_ok,
               double* se, double* p, double* ci_lo, double* ci_hi, double* t) {
    const int S = 3 * nq;
    int rc = orc_mm_pass(K, Xa, ya, na, Xb, yb, nb, sims, taus, draw_a, draw_b, nq, quantiles, point_stats,
                         point_betas_a, point_betas_b, NULL, NULL, NULL);                       /* :322 */
    if (rc != ORC_OK) return rc;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic)
    for (int64_t r = 0; r < reps; ++r) {                                                          /* :337-354 */
        double* Xsa = (double*)malloc(sizeof(double) * ((size_t)na * K + (size_t)nb * K + (size_t)na + (size_t)nb));
        double *Xsb = Xsa + (size_t)na * K, *ysa = Xsb + (size_t)nb * K, *ysb = ysa + na;
        for (int64_t i = 0; i < na; ++i) {
            const uint32_t k = idx_a[r * na + i];
            memcpy(Xsa + i * K, Xa + (size_t)k * K, sizeof(double) * (size_t)K); ysa[i] = ya[k];
        }
        for (int64_t i = 0; i < nb; ++i) {
            const uint32_t k = idx_b[r * nb + i];
            memcpy(Xsb + i * K, Xb + (size_t)k * K, sizeof(double) * (size_t)K); ysb[i] = yb[k];
        }
        const int prc = orc_mm_pass(K, Xsa, ysa, na, Xsb, ysb, nb, sims, taus + (size_t)(r + 1) * sims,
                                    draw_a + (size_t)(r + 1) * sims, draw_b + (size_t)(r + 1) * sims, nq, quantiles,
                                    rep_stats + (size_t)r * S, NULL, NULL, NULL, NULL, NULL);
        rep_status[r] = prc;
        if (prc != ORC_OK)
            for (int j = 0; j < S; ++j) rep_stats[(size_t)r * S + j] = NAN;
        free(Xsa);
    }
    if (se) orc_reduce(rep_stats, rep_status, reps, S, point_stats, n_ok, se, p, ci_lo, ci_hi, t);
    return ORC_OK;
}
