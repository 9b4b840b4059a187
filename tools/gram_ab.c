/* tools/gram_ab.c -- A/B timing of Gram-kernel variants through the C ABI, without Python: the same binary is run
 * against differently built libobboot.so (LD_LIBRARY_PATH picks the variant; tools/gram_ab.sh builds and runs them).
 *
 *   gcc -O2 -I include tools/gram_ab.c -L oaxaca_blinder_rs_b200/_lib -lobboot -lm -o tools/_ab/gram_ab
 *   LD_LIBRARY_PATH=tools/_ab/<variant> tools/_ab/gram_ab <n> <p> <reps> <runs> [<ncat>]
 *
 * Frame: n rows, p continuous predictors and ncat four-level categoricals (K = 1 + p + 3 ncat), weights, two groups,
 * 64-bit LCG data.  Prints, per run, the
 * CUDA-event time of the contraction kernel and of the whole call, the DMMA fraction of the 37.1 TFLOP/s issue peak,
 * and a bit-level checksum of the standard errors (all variants must print the same checksum: the sums are fixed by
 * the summation tree, not by the kernel's schedule). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "obboot.h"

static uint64_t lcg_state = 0x0B200ULL;
static inline double lcg_uniform(void) {
    lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(lcg_state >> 11) / 9007199254740992.0;
}

#define CHECK(call)                                                                                         \
    do {                                                                                                    \
        ob_status s__ = (call);                                                                             \
        if (s__ != OB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)s__, ob_last_error(ctx)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 10000000;
    const int p = argc > 2 ? atoi(argv[2]) : 50;
    const int64_t reps = argc > 3 ? atoll(argv[3]) : 2000;
    const int runs = argc > 4 ? atoi(argv[4]) : 3;
    const int ncat = argc > 5 ? atoi(argv[5]) : 0;
    ob_ctx* ctx = NULL;
    if (ob_ctx_create(0, &ctx) != OB_OK) { fprintf(stderr, "no B200 device: there is no CPU fallback\n"); return 2; }

    double** x = malloc(sizeof(double*) * (size_t)p);
    for (int j = 0; j < p; ++j) x[j] = malloc(sizeof(double) * (size_t)n);
    double* y = malloc(sizeof(double) * (size_t)n);
    double* w = malloc(sizeof(double) * (size_t)n);
    uint8_t* grp = malloc((size_t)n);
    int32_t** cat = malloc(sizeof(int32_t*) * (size_t)(ncat > 0 ? ncat : 1));
    int32_t* levels = malloc(sizeof(int32_t) * (size_t)(ncat > 0 ? ncat : 1));
    for (int q = 0; q < ncat; ++q) { cat[q] = malloc(sizeof(int32_t) * (size_t)n); levels[q] = 4; }
    for (int64_t i = 0; i < n; ++i) {
        grp[i] = lcg_uniform() < 0.5 ? 0 : 1;
        double acc = grp[i] == 0 ? 2.9 : 2.7;
        for (int j = 0; j < p; ++j) { x[j][i] = lcg_uniform() * 2.0 - 1.0 + (grp[i] == 0 ? 0.2 : 0.0); acc += 0.01 * (1 + j % 5) * x[j][i]; }
        for (int q = 0; q < ncat; ++q) { cat[q][i] = (int32_t)(lcg_uniform() * 4.0) & 3; acc += 0.05 * cat[q][i]; }
        w[i] = 0.5 + 2.5 * lcg_uniform();
        y[i] = acc + (lcg_uniform() - 0.5);
    }
    ob_frame_view f = {n, p, (const double* const*)x, ncat, (const int32_t* const*)cat, levels, y, w, grp};
    ob_design* d = NULL;
    CHECK(ob_design_pack(ctx, &f, &d));
    const int32_t K = p + 1 + 3 * ncat;
    const int32_t S = ob_num_stats(K, 0, NULL);
    ob_boot_opts o;
    memset(&o, 0, sizeof o);
    o.ref_kind = OB_REF_GROUP_B; o.reps = reps; o.seed = 1;
    double* buf = calloc((size_t)(6 * S + 3 * K), sizeof(double));
    ob_result r;
    memset(&r, 0, sizeof r);
    r.point_stats = buf; r.std_err = buf + S; r.p_value = buf + 2 * S; r.ci_lower = buf + 3 * S; r.ci_upper = buf + 4 * S;
    r.t_stat = buf + 5 * S; r.xa_mean = buf + 6 * S; r.xb_mean = buf + 6 * S + K; r.beta_star = buf + 6 * S + 2 * K;
    const double flop = 2.0 * (double)n * ((double)K * (K + 1) / 2 + K) * (double)reps;
    for (int it = 0; it < runs; ++it) {
        CHECK(ob_bootstrap_run(ctx, d, &o, &r));
        uint64_t h = 1469598103934665603ULL;
        for (int j = 0; j < S; ++j) { uint64_t b; memcpy(&b, &r.std_err[j], 8); h = (h ^ b) * 1099511628211ULL; }
        for (int j = 0; j < S; ++j) { uint64_t b; memcpy(&b, &r.point_stats[j], 8); h = (h ^ b) * 1099511628211ULL; }
        for (int j = 0; j < S; ++j) { uint64_t b; memcpy(&b, &r.ci_lower[j], 8); h = (h ^ b) * 1099511628211ULL; }
        printf("run %d gram_kernel_ms %.3f total_ms %.3f frac_of_37.1TF %.4f n_ok %lld se_hash %016llx\n", it, r.ms_gram_kernel,
               r.ms_total, flop / (r.ms_gram_kernel * 1e-3) / 37.1e12, (long long)r.n_ok, (unsigned long long)h);
    }
    ob_design_destroy(d);
    ob_ctx_destroy(ctx);
    return 0;
}
