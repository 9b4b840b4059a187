/* oracle/ob_oracle.h -- CPU restatement of the reference's bootstrap-inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oaxaca_blinder_rs_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: the reference (Rust; polars 0.44 + nalgebra 0.32.3 + rayon 1.11, none vendored,
 * no Cargo.lock, no cargo/rustc in this image) cannot be built here, so the oracle is pinned by
 * the reference's own known-answer unit tests and fixtures (tests/golden/ fixtures):
 * point estimates, OLS/Yun/decomposition/bootstrap_stats KATs are PINNED; the resampling RNG
 * stream and therefore every bootstrap SE/CI/p-value is "parity unpinned" in the reference
 * itself (sample_n_literal(.., seed=None), builder.rs:822-827) -- SE parity is GPU-vs-oracle
 * under an identical, harness-generated index stream.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/oaxaca_blinder/src/).
 */
#ifndef OB_ORACLE_H
#define OB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* error codes = OaxacaError variants (error.rs:6-19), numbered like include/obboot.h */
enum {
    ORC_OK = 0,
    ORC_ERR_POLARS = 1,
    ORC_ERR_COLUMN_NOT_FOUND = 2,
    ORC_ERR_INVALID_GROUP = 3,     /* empty group (builder.rs:431-435), negative weight (ols.rs:60-66) */
    ORC_ERR_NALGEBRA = 4,          /* Cholesky failed (ols.rs:107-111) */
    ORC_ERR_DIAGNOSTIC = 5,
    ORC_ERR_INSUFFICIENT_DATA = 6  /* n <= K (ols.rs:98-105) */
};

/* reference_coeffs (decomposition.rs:5-20): Neumark == Pooled, Cotton == Weighted */
enum { ORC_REF_GROUP_A = 0, ORC_REF_GROUP_B = 1, ORC_REF_POOLED = 2, ORC_REF_WEIGHTED = 3 };

/* Shape of one decomposition problem once the frame has been cleaned, dummy-expanded and
 * split (state at builder.rs:808).  Design columns: [__ob_intercept__, continuous.., dummies..]
 * (builder.rs:325-327). */
typedef struct {
    int32_t K;              /* design columns incl. intercept */
    int32_t n_cont;         /* continuous predictors; pooled indicator sits at column 1+n_cont (builder.rs:560-564 with :322-327) */
    int32_t ref_kind;       /* ORC_REF_* */
    int32_t n_norm;         /* entries of .normalize([...]) */
    const int32_t* norm_m;       /* [n_norm] level count m incl. base (category_counts, builder.rs:799) */
    const int32_t* norm_off;     /* [n_norm+1] offsets into norm_idx */
    const int32_t* norm_idx;     /* dummy column indices found by name prefix (normalization.rs:14-20) */
    const int32_t* norm_has_base;/* [n_norm] var present in base_categories (builder.rs:636-640) */
} orc_spec;

/* number of normalisation vars that contribute a base-category row */
int32_t orc_n_base(const orc_spec* s);
/* S = 5 + 2*(K + n_base): [explained, unexplained, endowments, coefficients, interaction,
 *                          detailed_explained[K+nb], detailed_unexplained[K+nb]] */
int32_t orc_n_stats(const orc_spec* s);

/* ols.rs:44-144 -- (W)LS by Cholesky on X'WX.  X row-major [n x K].  resid may be NULL.
 * precise != 0 accumulates X'WX / X'Wy in long double (checker mode); 0 = plain double
 * (reference-shaped arithmetic, used for the CPU baseline timing). */
int orc_ols(const double* y, const double* X, const double* w, int64_t n, int32_t K,
            int precise, double* beta, double* resid);

/* normalization.rs:5-51 -- Yun shift in place; base_coeff[v] = -mu (0 when the var has no dummies).
 * idx_shift_from: column index from which dummy indices are shifted by +1 (pooled design), or -1. */
void orc_yun(const orc_spec* s, double* beta, int32_t idx_shift_from, double* base_coeff);

/* decomposition.rs:56-70, :73-89, :92-122 */
void orc_two_fold(const double* xa, const double* xb, const double* ba, const double* bb,
                  const double* bs, int32_t K, double out[2]);
void orc_three_fold(const double* xa, const double* xb, const double* ba, const double* bb,
                    int32_t K, double out[3]);
void orc_detailed(const double* xa, const double* xb, const double* ba, const double* bb,
                  const double* bs, int32_t K, double* expl, double* unexpl);

/* inference.rs:4-34; out = {std_err, p_value, ci_lower, ci_upper} */
void orc_bootstrap_stats(const double* est, int64_t n, double out[4]);

/* math/rif.rs:14-88 */
void orc_rif(const double* y, int64_t n, double tau, double* rif_out);

typedef struct {
    double two_fold[2];
    double three_fold[3];
    double total_gap;
    double* det_expl;    /* [K+nb] */
    double* det_unexpl;  /* [K+nb] */
    double* xa_mean;     /* [K] */
    double* xb_mean;     /* [K] */
    double* beta_star;   /* [K] */
    double* beta_a;      /* [K] after Yun */
    double* beta_b;      /* [K] after Yun */
    double* resid_a;     /* [n_a] or NULL */
    double* resid_b;     /* [n_b] or NULL */
} orc_pass_out;

/* builder.rs:420-699 (OlsEstimator branch, estimation.rs:51-112) on already-split dense data. */
int orc_single_pass(const orc_spec* s,
                    const double* Xa, const double* ya, const double* wa, int64_t na,
                    const double* Xb, const double* yb, const double* wb, int64_t nb,
                    int precise, orc_pass_out* out);

typedef struct {
    /* point pass */
    orc_pass_out point;
    /* per replicate (any may be NULL) */
    double* rep_stats;    /* [reps x S] row-major, rows of failed replicates are NaN */
    int32_t* rep_status;  /* [reps] ORC_* code */
    double* rep_beta_a;   /* [reps x K] */
    double* rep_beta_b;   /* [reps x K] */
    double* rep_min_pivot;/* [reps] smallest Cholesky pivot relative to its diagonal (numerical-rank indicator) */
    /* reduction over the successful replicates, in replicate order (builder.rs:849-930) */
    int64_t n_ok;
    double* se;  double* p;  double* ci_lo;  double* ci_hi;  double* t;   /* [S] each */
} orc_run_out;

/* builder.rs:787-951.  idx_a [reps x na], idx_b [reps x nb] = explicit resample index stream;
 * NULL -> the oracle's own xoshiro stream keyed by (seed, replicate, group) (the reference's
 * stream is unseeded, so no particular stream is "the" reference stream).
 * Replicates: gather rows like sample_n_literal + vstack (builder.rs:822-829) then the single pass.
 * nthreads: OpenMP threads over replicates (rayon par_iter, builder.rs:816-817). */
int orc_run(const orc_spec* s,
            const double* Xa, const double* ya, const double* wa, int64_t na,
            const double* Xb, const double* yb, const double* wb, int64_t nb,
            int64_t reps, const uint32_t* idx_a, const uint32_t* idx_b, uint64_t seed,
            int nthreads, int precise, orc_run_out* out);

/* the oracle's own index stream, exposed so tests can feed the same stream to the GPU path */
void orc_fill_indices(uint64_t seed, int64_t rep, int32_t group, int64_t n, uint32_t* idx);

/* reduction only (used by multi-rank tests after gathering replicate rows) */
void orc_reduce(const double* rep_stats, const int32_t* rep_status, int64_t reps, int32_t S,
                const double* point_stats, int64_t* n_ok,
                double* se, double* p, double* ci_lo, double* ci_hi, double* t);

/* numerical-rank indicator of the Cholesky factorisations since the last reset (thread-local) */
double orc_last_min_pivot(void);
void orc_reset_min_pivot(void);

/* ---- Heckman two-step replicate (ob_oracle_heckman.c).  X [n x K] outcome design incl. intercept, Z [n x K1]
 * selection design (intercept first), s [n] selection outcome (0 / 1), all rows of the group; the outcome equation uses
 * the rows with s == 1.  Statistics: S = 5 + 2 (K + 1) + K1 =
 * [explained, unexplained, endowments, coefficients, interaction, det_expl[K+1], det_unexpl[K+1] (IMR last), selection[K1]]. */
int orc_probit(const double* y, const double* X, int64_t n, int32_t k, int32_t max_iter, double tol,
               double* beta, int32_t* converged, int32_t* iterations);                     /* math/probit.rs:25-175 */
int32_t orc_heckman_n_stats(int32_t K, int32_t K1);
int orc_heckman_pass(int32_t K, int32_t K1, int32_t ref_kind,
                     const double* Xa, const double* ya, const double* Za, const double* sa, int64_t na,
                     const double* Xb, const double* yb, const double* Zb, const double* sb, int64_t nb,
                     int precise, double* stats, double* beta_a, double* beta_b, double* gamma_a, double* gamma_b,
                     double* total_gap);                                                    /* builder.rs:420-699, estimation.rs:114-269 */
int orc_heckman_run(int32_t K, int32_t K1, int32_t ref_kind,
                    const double* Xa, const double* ya, const double* Za, const double* sa, int64_t na,
                    const double* Xb, const double* yb, const double* Zb, const double* sb, int64_t nb,
                    int64_t reps, const uint32_t* idx_a, const uint32_t* idx_b, int nthreads, int precise,
                    double* point_stats, double* point_beta_a, double* point_beta_b, double* point_gamma_a, double* point_gamma_b,
                    double* total_gap, double* rep_stats, int32_t* rep_status, double* rep_gamma_a,
                    int64_t* n_ok, double* se, double* p, double* ci_lo, double* ci_hi, double* t);

/* ---- Machado-Mata quantile decomposition (ob_oracle_mm.c) ----
 * orc_qr: math/quantile_regression.rs:22-135 (the LP's vertex; c = optional multiplicities, NULL = 1);
 * info [3] = {interior-point iterations, status 0 vertex verified / 1 interior-point solution only / 2 failed, rows used by the polish}.
 * orc_mm_pass: run_single_pass, quantile_decomposition.rs:173-279; orc_mm_run: run(), :281-421.  Statistics per pass:
 * [gap, characteristics, coefficients] per target quantile. */
int orc_qr(const double* X, const double* y, const double* c, int64_t n, int32_t K, double tau, double* beta, int32_t* info);
int orc_mm_pass(int32_t K, const double* Xa, const double* ya, int64_t na, const double* Xb, const double* yb, int64_t nb,
                int32_t sims, const double* taus, const uint32_t* draw_a, const uint32_t* draw_b,
                int32_t nq, const double* quantiles, double* stats,
                double* betas_a, double* betas_b, int32_t* qr_status_a, int32_t* qr_status_b, int32_t* nsucc_out);
int orc_mm_run(int32_t K, const double* Xa, const double* ya, int64_t na, const double* Xb, const double* yb, int64_t nb,
               int32_t sims, int32_t nq, const double* quantiles, int64_t reps,
               const uint32_t* idx_a, const uint32_t* idx_b, const double* taus, const uint32_t* draw_a, const uint32_t* draw_b,
               int nthreads, double* point_stats, double* point_betas_a, double* point_betas_b,
               double* rep_stats, int32_t* rep_status, int64_t* n_ok,
               double* se, double* p, double* ci_lo, double* ci_hi, double* t);

/* flatten a pass into the S-vector layout */
void orc_pass_to_stats(const orc_spec* s, const orc_pass_out* p, double* stats);

#ifdef __cplusplus
}
#endif
#endif
