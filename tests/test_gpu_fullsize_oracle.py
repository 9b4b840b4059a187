"""GPU vs the oracle AT THE BASELINE SHAPES (BASELINE.json configs[1..4]), replicate by replicate under an explicit
resample index stream -- the split-n summation over millions of rows per group is exactly where order effects live.

  config 2  n = 1M,   K = 21, OLS, three-fold                      4 replicates
  config 3  n = 10M,  K = 51, WLS + Yun on both categoricals       4 replicates   (the headline shape)
  config 4  n = 5M,   K = 31, RIF outcome (tau = 0.9), OLS         3 replicates   apply_rif -> bootstrap vs the oracle
                                                                                   run on orc.rif(y)
  config 5  n = 1e8 (host RAM >= 160 GB, else 2e7), K = 17, OLS    2 replicates

Asserted: replicate status equal; per-replicate coefficients, statistics and the point estimates within 1e-10
ELEMENTWISE relative (tests/helpers.relerr; reference ols.rs:44-144, builder.rs:822-839); multiplicities of the
index stream bit-exact against np.bincount.  The oracle accumulates X'WX in double-double, so it is the more
accurate side.
"""
import os

import numpy as np
import pytest

from helpers import relerr, relerr_to_scale

pytestmark = pytest.mark.gpu

RTOL = 1e-10      # north_star


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def _fullsize_case(orc, n, n_cont, cats, weights, normalize, reps, ref_kind, rif_tau=None, seed=4242):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(n, n_cont, cat_levels=cats, weights=weights)
    norm = synth.norm_spec(d) if normalize else []
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    na, nb = len(ya), len(yb)
    ia, ib = orc.index_stream(seed, reps, 0, na), orc.index_stream(seed, reps, 1, nb)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    del d
    try:
        assert (des.n_a, des.n_b) == (na, nb)
        if rif_tau is not None:            # decompose_quantile (builder.rs:721-737): RIF once per group, then run()
            des.apply_rif(rif_tau)
            ya_o, yb_o = orc.rif(ya, rif_tau), orc.rif(yb, rif_tau)
            _, ga, _, _, gb, _ = des.download()
            assert relerr(ga, ya_o) <= RTOL and relerr(gb, yb_o) <= RTOL
        else:
            ya_o, yb_o = ya, yb
        gpu = ob.bootstrap(des, reps, ref_kind=ref_kind, norm=[ob.NormVar(m, i) for m, i in norm], idx_a=ia, idx_b=ib,
                           want_rep=True)
        # multiplicities of the same stream, straight from the production histogram kernel
        if n <= 10_000_000:
            ca, fa = des.debug_counts_from_indices(ia[:2], 0)
            assert fa == 0 and all(np.array_equal(ca[r], np.bincount(ia[r], minlength=na)) for r in range(2))
    finally:
        des.close()
        ctx.close()
    spec = orc.Spec(K=Xa.shape[1], n_cont=n_cont, ref_kind=ref_kind, norm=[orc.NormVar(m, i) for m, i in norm])
    ref = orc.run(spec, Xa, ya_o, wa, Xb, yb_o, wb, reps, ia, ib, nthreads=reps + 1, precise=True)
    return gpu, ref


def _assert_parity(gpu, ref):
    p = ref["point"]
    assert np.array_equal(gpu["rep_status"], ref["rep_status"]) and gpu["n_ok"] == ref["n_ok"]
    assert abs(gpu["total_gap"] - p["total_gap"]) <= RTOL * abs(p["total_gap"])
    errs = {}
    for k_gpu, k_ref in (("point_stats", "stats"), ("xa_mean", "xa_mean"), ("xb_mean", "xb_mean"),
                         ("beta_star", "beta_star"), ("beta_a", "beta_a"), ("beta_b", "beta_b")):
        errs[k_gpu] = relerr(gpu[k_gpu], p[k_ref])
    errs["rep_stats"] = relerr(gpu["rep_stats"], ref["rep_stats"])
    errs["rep_beta_a"] = relerr(gpu["rep_beta_a"], ref["rep_beta_a"])
    errs["rep_beta_b"] = relerr(gpu["rep_beta_b"], ref["rep_beta_b"])
    errs["std_err"] = relerr(gpu["std_err"], ref["se"])
    errs["ci_lower"] = relerr(gpu["ci_lower"], ref["ci_lo"])
    errs["ci_upper"] = relerr(gpu["ci_upper"], ref["ci_hi"])
    yscale = max(1.0, float(np.max(np.abs(p["resid_b"]))))
    errs["residuals_b"] = relerr_to_scale(gpu["residuals_b"], p["resid_b"], yscale)
    print("elementwise relative errors vs the oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v <= RTOL}
    assert not bad, bad
    return errs


def test_config3_n10M_k50_wls_yun_vs_oracle(orc):
    gpu, ref = _fullsize_case(orc, 10_000_000, 44, (4, 4), True, True, reps=4, ref_kind=3)   # Cotton: omega from the resampled weights
    _assert_parity(gpu, ref)


def test_config2_n1M_k20_threefold_vs_oracle(orc):
    gpu, ref = _fullsize_case(orc, 1_000_000, 20, (), False, False, reps=4, ref_kind=2)      # pooled beta*
    _assert_parity(gpu, ref)


def test_config4_n5M_k30_rif_vs_oracle(orc):
    gpu, ref = _fullsize_case(orc, 5_000_000, 30, (), False, False, reps=3, ref_kind=0, rif_tau=0.9)
    _assert_parity(gpu, ref)


def test_config5_shape_k16_vs_oracle(orc):
    n = 100_000_000 if (_host_ram_gb() >= 160 and os.environ.get("OB_FULLSIZE_SMALL") != "1") else 20_000_000
    print("config-5 shape at n =", n)
    gpu, ref = _fullsize_case(orc, n, 16, (), False, False, reps=2, ref_kind=1)
    _assert_parity(gpu, ref)
