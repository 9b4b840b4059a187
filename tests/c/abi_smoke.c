/* tests/c/abi_smoke.c -- the drop-in boundary used from plain C (what a build.rs-linked extern "C" block does):
 * no Python, no torch, nothing but include/obboot.h and libobboot.so.
 *
 *   gcc -O2 -I include tests/c/abi_smoke.c -L oaxaca_blinder_rs_b200/_lib -lobboot -Wl,-rpath,... -lm -o abi_smoke
 *   ./abi_smoke <n> <reps> <seed>        prints one line per number, "%.17g"
 *
 * Frame: a 64-bit LCG anyone can restate (tests/test_gpu_cabi.py does, in numpy): group, two continuous predictors, one
 * categorical with 3 levels, weights, outcome.  Runs ob_design_pack_async -> ob_bootstrap_run (native stream, pooled
 * beta*, Yun on the categorical) with page-locked columns, then once more through ob_design_pack for comparison, then
 * ob_mm_run (Machado-Mata, native streams) on the unweighted pack of the same frame. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "obboot.h"

static uint64_t lcg_state;
static double lcg_uniform(void) {               /* Knuth MMIX LCG, top 53 bits */
    lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(lcg_state >> 11) / 9007199254740992.0;
}

#define CHECK(call)                                                                             \
    do {                                                                                        \
        ob_status s__ = (call);                                                                 \
        if (s__ != OB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)s__, ob_last_error(ctx)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 20000;
    const int64_t reps = argc > 2 ? atoll(argv[2]) : 64;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
    ob_ctx* ctx = NULL;
    if (ob_ctx_create(0, &ctx) != OB_OK) { fprintf(stderr, "no B200 device: there is no CPU fallback\n"); return 2; }

    double *x0, *x1, *y, *w; int32_t* cat; uint8_t* grp;
    void* p;
    CHECK(ob_host_alloc(sizeof(double) * n, &p)); x0 = p;
    CHECK(ob_host_alloc(sizeof(double) * n, &p)); x1 = p;
    CHECK(ob_host_alloc(sizeof(double) * n, &p)); y = p;
    CHECK(ob_host_alloc(sizeof(double) * n, &p)); w = p;
    CHECK(ob_host_alloc(sizeof(int32_t) * n, &p)); cat = p;
    CHECK(ob_host_alloc(n, &p)); grp = p;
    lcg_state = 0x0B200ULL;
    for (int64_t i = 0; i < n; ++i) {
        const double ug = lcg_uniform(), u0 = lcg_uniform(), u1 = lcg_uniform(), uc = lcg_uniform(), uw = lcg_uniform(), ue = lcg_uniform();
        grp[i] = ug < 0.5 ? 0 : 1;
        x0[i] = 8.0 + 12.0 * u0 + (grp[i] == 0 ? 0.5 : 0.0);
        x1[i] = 40.0 * u1;
        cat[i] = uc < 0.4 ? 0 : (uc < 0.75 ? 1 : 2);
        w[i] = 0.5 + 2.5 * uw;
        y[i] = (grp[i] == 0 ? 2.9 : 2.7) + 0.08 * x0[i] + 0.01 * x1[i] + 0.1 * cat[i] + (ue - 0.5);
    }
    const double* cont[2] = {x0, x1};
    const int32_t* cats[1] = {cat};
    const int32_t levels[1] = {3};
    ob_frame_view f = {n, 2, cont, 1, cats, levels, y, w, grp};

    const int32_t K = 1 + 2 + 2;
    const int32_t norm_m[1] = {3}, norm_off[2] = {0, 2}, norm_idx[2] = {3, 4}, norm_hb[1] = {1};
    const int32_t S = ob_num_stats(K, 1, norm_hb);
    ob_boot_opts o;
    memset(&o, 0, sizeof o);
    o.ref_kind = OB_REF_POOLED; o.n_norm = 1; o.norm_m = norm_m; o.norm_off = norm_off; o.norm_idx = norm_idx; o.norm_has_base = norm_hb;
    o.reps = reps; o.seed = seed;

    double* out[2];
    for (int pass = 0; pass < 2; ++pass) {
        ob_design* d = NULL;
        if (pass == 0) CHECK(ob_design_pack_async(ctx, &f, &d)); else CHECK(ob_design_pack(ctx, &f, &d));
        int64_t na = 0, nb = 0;
        CHECK(ob_design_shape(d, &na, &nb, NULL, NULL));
        ob_result r;
        memset(&r, 0, sizeof r);
        double* buf = calloc((size_t)(6 * S + 3 * K), sizeof(double));
        r.point_stats = buf; r.std_err = buf + S; r.p_value = buf + 2 * S; r.ci_lower = buf + 3 * S; r.ci_upper = buf + 4 * S;
        r.t_stat = buf + 5 * S; r.xa_mean = buf + 6 * S; r.xb_mean = buf + 6 * S + K; r.beta_star = buf + 6 * S + 2 * K;
        void* res = NULL;
        CHECK(ob_host_alloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1), &res));
        r.residuals_b = res;
        CHECK(ob_bootstrap_run(ctx, d, &o, &r));
        if (pass == 0) {
            printf("n_a %lld\nn_b %lld\nS %d\nn_ok %lld\ntotal_gap %.17g\n", (long long)na, (long long)nb, S, (long long)r.n_ok, r.total_gap);
            for (int j = 0; j < S; ++j) printf("point %.17g\nse %.17g\nci_lo %.17g\nci_hi %.17g\n", r.point_stats[j], r.std_err[j], r.ci_lower[j], r.ci_upper[j]);
            for (int j = 0; j < K; ++j) printf("beta_star %.17g\n", r.beta_star[j]);
            double rs = 0.0;
            for (int64_t i = 0; i < nb; ++i) rs += fabs(r.residuals_b[i]);
            printf("resid_abs_sum %.17g\n", rs);
        }
        out[pass] = buf;
        ob_host_free(res);
        ob_design_destroy(d);
    }
    const int same = memcmp(out[0], out[1], sizeof(double) * (size_t)(6 * S + 3 * K)) == 0;   /* async pack == pack, bit for bit */
    printf("async_equals_sync %d\n", same);
    {   /* Machado-Mata (QuantileDecompositionBuilder::run): the same frame without weights, native streams */
        ob_frame_view fu = f;
        fu.weights = NULL;
        ob_design* d = NULL;
        CHECK(ob_design_pack(ctx, &fu, &d));
        const double quantiles[3] = {0.25, 0.5, 0.75};
        ob_mm_opts mo;
        memset(&mo, 0, sizeof mo);
        mo.simulations = 24; mo.n_quantiles = 3; mo.quantiles = quantiles; mo.reps = 5; mo.seed = seed;
        double mm[6 * 9];
        ob_mm_result mr;
        memset(&mr, 0, sizeof mr);
        mr.point_stats = mm; mr.std_err = mm + 9; mr.p_value = mm + 18; mr.ci_lower = mm + 27; mr.ci_upper = mm + 36; mr.t_stat = mm + 45;
        CHECK(ob_mm_run(ctx, d, &mo, &mr));
        printf("mm_n_ok %lld\nmm_qr_total %lld\nmm_qr_failed %lld\n", (long long)mr.n_ok, (long long)mr.qr_total, (long long)mr.qr_failed);
        for (int j = 0; j < 9; ++j) printf("mm_point %.17g\nmm_se %.17g\nmm_ci_lo %.17g\nmm_ci_hi %.17g\n", mm[j], mm[9 + j], mm[27 + j], mm[36 + j]);
        ob_design_destroy(d);
    }
    free(out[0]); free(out[1]);
    ob_host_free(x0); ob_host_free(x1); ob_host_free(y); ob_host_free(w); ob_host_free(cat); ob_host_free(grp);
    ob_ctx_destroy(ctx);
    return same ? 0 : 3;
}
