"""ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

REF_GROUP_A, REF_GROUP_B, REF_POOLED, REF_WEIGHTED = 0, 1, 2, 3
ERR_NAMES = {0: "Ok", 1: "PolarsError", 2: "ColumnNotFound", 3: "InvalidGroupVariable",
             4: "NalgebraError", 5: "DiagnosticError", 6: "InsufficientData"}


def build(force: bool = False) -> str:
    """Compile the oracle (gcc, a second or two).  Building the checker is not using it."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("ob_oracle.c", "ob_oracle_heckman.c", "ob_oracle_mm.c", "ob_oracle.h"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


class _Spec(C.Structure):
    _fields_ = [("K", C.c_int32), ("n_cont", C.c_int32), ("ref_kind", C.c_int32), ("n_norm", C.c_int32),
                ("norm_m", C.POINTER(C.c_int32)), ("norm_off", C.POINTER(C.c_int32)),
                ("norm_idx", C.POINTER(C.c_int32)), ("norm_has_base", C.POINTER(C.c_int32))]


_DP = C.POINTER(C.c_double)


class _PassOut(C.Structure):
    _fields_ = [("two_fold", C.c_double * 2), ("three_fold", C.c_double * 3), ("total_gap", C.c_double),
                ("det_expl", _DP), ("det_unexpl", _DP), ("xa_mean", _DP), ("xb_mean", _DP),
                ("beta_star", _DP), ("beta_a", _DP), ("beta_b", _DP), ("resid_a", _DP), ("resid_b", _DP)]


class _RunOut(C.Structure):
    _fields_ = [("point", _PassOut), ("rep_stats", _DP), ("rep_status", C.POINTER(C.c_int32)),
                ("rep_beta_a", _DP), ("rep_beta_b", _DP), ("rep_min_pivot", _DP), ("n_ok", C.c_int64),
                ("se", _DP), ("p", _DP), ("ci_lo", _DP), ("ci_hi", _DP), ("t", _DP)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_n_base.restype = C.c_int32
        L.orc_n_stats.restype = C.c_int32
        L.orc_ols.restype = C.c_int
        L.orc_ols.argtypes = [_DP, _DP, _DP, C.c_int64, C.c_int32, C.c_int, _DP, _DP]
        L.orc_single_pass.restype = C.c_int
        L.orc_run.restype = C.c_int
        L.orc_bootstrap_stats.argtypes = [_DP, C.c_int64, _DP]
        L.orc_rif.argtypes = [_DP, C.c_int64, C.c_double, _DP]
        L.orc_fill_indices.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_int64, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_DP)


def _f64(a, shape=None) -> Optional[np.ndarray]:
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


@dataclass
class NormVar:
    """One entry of .normalize([...]): level count m (incl. base), dummy column indices in the K-column
    design (found by name prefix, normalization.rs:14-20) and whether a base row is emitted."""
    m: int
    idx: Sequence[int]
    has_base: bool = True


@dataclass
class Spec:
    K: int
    n_cont: int
    ref_kind: int = REF_GROUP_A
    norm: List[NormVar] = field(default_factory=list)

    def _c(self) -> Tuple[_Spec, list]:
        m = np.array([v.m for v in self.norm] + [0], dtype=np.int32)
        off = np.zeros(len(self.norm) + 1, dtype=np.int32)
        for i, v in enumerate(self.norm):
            off[i + 1] = off[i] + len(v.idx)
        idx = np.array([j for v in self.norm for j in v.idx] + [0], dtype=np.int32)
        hb = np.array([int(v.has_base) for v in self.norm] + [0], dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        s = _Spec(self.K, self.n_cont, self.ref_kind, len(self.norm), m.ctypes.data_as(ip),
                  off.ctypes.data_as(ip), idx.ctypes.data_as(ip), hb.ctypes.data_as(ip))
        return s, [m, off, idx, hb]

    @property
    def n_base(self) -> int:
        return sum(1 for v in self.norm if v.has_base)

    @property
    def n_stats(self) -> int:
        return 5 + 2 * (self.K + self.n_base)


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle: {ERR_NAMES.get(code, code)}")
        self.code = code


def ols(y, X, w=None, precise=True, want_resid=True):
    """ols.rs:44-144.  Returns (beta, residuals)."""
    X = _f64(X)
    n, K = X.shape
    y = _f64(y, (n,))
    w = _f64(w, (n,)) if w is not None else None
    beta = np.empty(K)
    resid = np.empty(n) if want_resid else None
    rc = lib().orc_ols(_dp(y), _dp(X), _dp(w), n, K, int(precise), _dp(beta), _dp(resid))
    if rc:
        raise OracleError(rc)
    return beta, resid


def yun(spec: Spec, beta, idx_shift_from: int = -1):
    beta = np.array(beta, dtype=np.float64)
    base = np.zeros(max(1, len(spec.norm)))
    cs, keep = spec._c()
    lib().orc_yun(C.byref(cs), _dp(beta), C.c_int32(idx_shift_from), _dp(base))
    return beta, base[:len(spec.norm)]


def bootstrap_stats(est):
    est = _f64(est)
    out = np.empty(4)
    lib().orc_bootstrap_stats(_dp(est), est.size, _dp(out))
    return {"se": out[0], "p": out[1], "ci_lo": out[2], "ci_hi": out[3]}


def rif(y, tau: float):
    y = _f64(y)
    out = np.empty_like(y)
    lib().orc_rif(_dp(y), y.size, float(tau), _dp(out))
    return out


def fill_indices(seed: int, rep: int, group: int, n: int) -> np.ndarray:
    idx = np.empty(n, dtype=np.uint32)
    lib().orc_fill_indices(seed, rep, group, n, idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return idx


def index_stream(seed: int, reps: int, group: int, n: int) -> np.ndarray:
    out = np.empty((reps, n), dtype=np.uint32)
    for b in range(reps):
        out[b] = fill_indices(seed, b, group, n)
    return out


def _alloc_pass(spec: Spec, na: int, nb: int, want_resid: bool):
    D = spec.K + spec.n_base
    arrs = dict(det_expl=np.empty(D), det_unexpl=np.empty(D), xa_mean=np.empty(spec.K),
                xb_mean=np.empty(spec.K), beta_star=np.empty(spec.K), beta_a=np.empty(spec.K),
                beta_b=np.empty(spec.K),
                resid_a=np.empty(na) if want_resid else None,
                resid_b=np.empty(nb) if want_resid else None)
    po = _PassOut()
    for k, v in arrs.items():
        setattr(po, k, _dp(v))
    return po, arrs


def _pass_dict(po: _PassOut, arrs: dict) -> dict:
    d = dict(arrs)
    d["two_fold"] = np.array(list(po.two_fold))
    d["three_fold"] = np.array(list(po.three_fold))
    d["total_gap"] = float(po.total_gap)
    d["stats"] = np.concatenate([d["two_fold"], d["three_fold"], d["det_expl"], d["det_unexpl"]])
    return d


def _prep(spec, Xa, ya, wa, Xb, yb, wb):
    Xa, Xb = _f64(Xa), _f64(Xb)
    na, nb = Xa.shape[0], Xb.shape[0]
    assert Xa.shape[1] == spec.K and Xb.shape[1] == spec.K
    ya, yb = _f64(ya, (na,)), _f64(yb, (nb,))
    wa = _f64(wa, (na,)) if wa is not None else None
    wb = _f64(wb, (nb,)) if wb is not None else None
    return Xa, ya, wa, na, Xb, yb, wb, nb


def single_pass(spec: Spec, Xa, ya, wa, Xb, yb, wb, precise=True, want_resid=True) -> dict:
    """builder.rs:420-699 on dense, already split data."""
    Xa, ya, wa, na, Xb, yb, wb, nb = _prep(spec, Xa, ya, wa, Xb, yb, wb)
    cs, keep = spec._c()
    po, arrs = _alloc_pass(spec, na, nb, want_resid)
    rc = lib().orc_single_pass(C.byref(cs), _dp(Xa), _dp(ya), _dp(wa), C.c_int64(na),
                               _dp(Xb), _dp(yb), _dp(wb), C.c_int64(nb), int(precise), C.byref(po))
    if rc:
        raise OracleError(rc)
    return _pass_dict(po, arrs)


def run(spec: Spec, Xa, ya, wa, Xb, yb, wb, reps: int, idx_a=None, idx_b=None, seed: int = 0,
        nthreads: int = 1, precise=True, want_rep=True) -> dict:
    """builder.rs:787-951: point pass + `reps` replicates + bootstrap_stats reduction."""
    Xa, ya, wa, na, Xb, yb, wb, nb = _prep(spec, Xa, ya, wa, Xb, yb, wb)
    cs, keep = spec._c()
    S, K = spec.n_stats, spec.K
    ro = _RunOut()
    po, arrs = _alloc_pass(spec, na, nb, True)
    ro.point = po
    R = max(reps, 1)
    rep_stats = np.full((R, S), np.nan)
    rep_status = np.zeros(R, dtype=np.int32)
    rep_beta_a = np.full((R, K), np.nan) if want_rep else None
    rep_beta_b = np.full((R, K), np.nan) if want_rep else None
    red = {k: np.empty(S) for k in ("se", "p", "ci_lo", "ci_hi", "t")}
    ro.rep_stats, ro.rep_status = _dp(rep_stats), rep_status.ctypes.data_as(C.POINTER(C.c_int32))
    ro.rep_beta_a, ro.rep_beta_b = _dp(rep_beta_a), _dp(rep_beta_b)
    rep_min_pivot = np.ones(R)
    ro.rep_min_pivot = _dp(rep_min_pivot)
    for k, v in red.items():
        setattr(ro, k, _dp(v))
    u32p = C.POINTER(C.c_uint32)
    if idx_a is not None:
        idx_a = np.ascontiguousarray(idx_a, dtype=np.uint32)
        assert idx_a.shape == (reps, na)
    if idx_b is not None:
        idx_b = np.ascontiguousarray(idx_b, dtype=np.uint32)
        assert idx_b.shape == (reps, nb)
    rc = lib().orc_run(C.byref(cs), _dp(Xa), _dp(ya), _dp(wa), C.c_int64(na), _dp(Xb), _dp(yb), _dp(wb),
                       C.c_int64(nb), C.c_int64(reps),
                       None if idx_a is None else idx_a.ctypes.data_as(u32p),
                       None if idx_b is None else idx_b.ctypes.data_as(u32p),
                       C.c_uint64(seed), int(nthreads), int(precise), C.byref(ro))
    if rc:
        raise OracleError(rc)
    out = {"point": _pass_dict(ro.point, arrs), "n_ok": int(ro.n_ok), "rep_stats": rep_stats[:reps],
           "rep_status": rep_status[:reps], "rep_min_pivot": rep_min_pivot[:reps]}
    if want_rep:
        out["rep_beta_a"], out["rep_beta_b"] = rep_beta_a[:reps], rep_beta_b[:reps]
    out.update(red)
    return out


def reduce(rep_stats, rep_status, point_stats) -> dict:
    rep_stats = _f64(rep_stats)
    reps, S = rep_stats.shape
    rep_status = np.ascontiguousarray(rep_status, dtype=np.int32)
    point_stats = _f64(point_stats, (S,))
    red = {k: np.empty(S) for k in ("se", "p", "ci_lo", "ci_hi", "t")}
    n_ok = C.c_int64(0)
    lib().orc_reduce(_dp(rep_stats), rep_status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(reps),
                     C.c_int32(S), _dp(point_stats), C.byref(n_ok), *[_dp(red[k]) for k in
                                                                      ("se", "p", "ci_lo", "ci_hi", "t")])
    red["n_ok"] = int(n_ok.value)
    return red


# ---- Heckman two-step replicate (ob_oracle_heckman.c) -------------------------------------------------------------
def probit(y, X, max_iter: int = 100, tol: float = 1e-6):
    """math/probit.rs:25-175 -> (beta, converged, iterations)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    beta = np.empty(X.shape[1])
    conv, it = C.c_int32(), C.c_int32()
    rc = lib().orc_probit(_dp(y), _dp(X), C.c_int64(X.shape[0]), C.c_int32(X.shape[1]), C.c_int32(max_iter), C.c_double(tol),
                          _dp(beta), C.byref(conv), C.byref(it))
    if rc != 0:
        raise RuntimeError(ERR_NAMES.get(rc, str(rc)))
    return beta, bool(conv.value), it.value


def heckman_run(ref_kind, Xa, ya, Za, sa, Xb, yb, Zb, sb, reps, idx_a, idx_b, nthreads=4, precise=True):
    """builder.rs:787-951 with the HeckmanEstimator (estimation.rs:114-269) under an explicit index stream."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Xa, ya, Za, sa, Xb, yb, Zb, sb)]
    Xa, ya, Za, sa, Xb, yb, Zb, sb = arrs
    K, K1, na, nb = Xa.shape[1], Za.shape[1], Xa.shape[0], Xb.shape[0]
    L = lib()
    L.orc_heckman_n_stats.restype = C.c_int32
    S = L.orc_heckman_n_stats(C.c_int32(K), C.c_int32(K1))
    idx_a = np.ascontiguousarray(idx_a if reps else np.zeros((0, na)), dtype=np.uint32)
    idx_b = np.ascontiguousarray(idx_b if reps else np.zeros((0, nb)), dtype=np.uint32)
    out = dict(point_stats=np.empty(S), beta_a=np.empty(K + 1), beta_b=np.empty(K + 1), gamma_a=np.empty(K1), gamma_b=np.empty(K1),
               rep_stats=np.empty((max(reps, 1), S)), rep_status=np.zeros(max(reps, 1), dtype=np.int32),
               rep_gamma_a=np.empty((max(reps, 1), K1)), se=np.empty(S), p=np.empty(S), ci_lo=np.empty(S), ci_hi=np.empty(S), t=np.empty(S))
    gap, nok = C.c_double(), C.c_int64()
    U32 = C.POINTER(C.c_uint32)
    rc = L.orc_heckman_run(C.c_int32(K), C.c_int32(K1), C.c_int32(ref_kind), _dp(Xa), _dp(ya), _dp(Za), _dp(sa), C.c_int64(na),
                           _dp(Xb), _dp(yb), _dp(Zb), _dp(sb), C.c_int64(nb), C.c_int64(reps), idx_a.ctypes.data_as(U32),
                           idx_b.ctypes.data_as(U32), C.c_int(nthreads), C.c_int(int(precise)),
                           _dp(out["point_stats"]), _dp(out["beta_a"]), _dp(out["beta_b"]), _dp(out["gamma_a"]), _dp(out["gamma_b"]),
                           C.byref(gap), _dp(out["rep_stats"]), out["rep_status"].ctypes.data_as(C.POINTER(C.c_int32)),
                           _dp(out["rep_gamma_a"]), C.byref(nok), _dp(out["se"]), _dp(out["p"]), _dp(out["ci_lo"]), _dp(out["ci_hi"]),
                           _dp(out["t"]))
    if rc != 0:
        raise RuntimeError(ERR_NAMES.get(rc, str(rc)))
    for k in ("rep_stats", "rep_status", "rep_gamma_a"):
        out[k] = out[k][:reps]
    out.update(total_gap=gap.value, n_ok=int(nok.value), S=S)
    return out


QR_VERTEX, QR_APPROX, QR_FAILED = 0, 1, 2


def qr(X, y, tau: float, c=None):
    """math/quantile_regression.rs:22-135 -> (beta, info) with info = dict(iters, status, ncand); the LP's vertex."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    cc = None if c is None else np.ascontiguousarray(c, dtype=np.float64)
    beta = np.full(X.shape[1], np.nan)
    info = np.zeros(3, dtype=np.int32)
    lib().orc_qr(_dp(X), _dp(y), _dp(cc), C.c_int64(X.shape[0]), C.c_int32(X.shape[1]), C.c_double(tau), _dp(beta),
                 info.ctypes.data_as(C.POINTER(C.c_int32)))
    return beta, dict(iters=int(info[0]), status=int(info[1]), ncand=int(info[2]))


def mm_pass(Xa, ya, Xb, yb, taus, draw_a, draw_b, quantiles):
    """run_single_pass, quantile_decomposition.rs:173-279 -> dict(stats [nq x 3], betas_a, betas_b, status_a, status_b, nsucc)."""
    Xa, ya, Xb, yb = [np.ascontiguousarray(a, dtype=np.float64) for a in (Xa, ya, Xb, yb)]
    taus = np.ascontiguousarray(taus, dtype=np.float64)
    quantiles = np.ascontiguousarray(quantiles, dtype=np.float64)
    da = np.ascontiguousarray(draw_a, dtype=np.uint32)
    db = np.ascontiguousarray(draw_b, dtype=np.uint32)
    K, sims, nq = Xa.shape[1], len(taus), len(quantiles)
    stats = np.full(3 * nq, np.nan)
    ba, bb = np.empty((sims, K)), np.empty((sims, K))
    sa, sb = np.zeros(sims, dtype=np.int32), np.zeros(sims, dtype=np.int32)
    ns = C.c_int32()
    U32, I32 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L = lib()
    L.orc_mm_pass.restype = C.c_int
    rc = L.orc_mm_pass(C.c_int32(K), _dp(Xa), _dp(ya), C.c_int64(Xa.shape[0]), _dp(Xb), _dp(yb), C.c_int64(Xb.shape[0]),
                       C.c_int32(sims), _dp(taus), da.ctypes.data_as(U32), db.ctypes.data_as(U32), C.c_int32(nq), _dp(quantiles),
                       _dp(stats), _dp(ba), _dp(bb), sa.ctypes.data_as(I32), sb.ctypes.data_as(I32), C.byref(ns))
    return dict(rc=rc, stats=stats.reshape(nq, 3), betas_a=ba, betas_b=bb, status_a=sa, status_b=sb, nsucc=ns.value)


def mm_run(Xa, ya, Xb, yb, sims, quantiles, reps, idx_a, idx_b, taus, draw_a, draw_b, nthreads=8):
    """QuantileDecompositionBuilder::run, quantile_decomposition.rs:281-421, under explicit streams (see ob_oracle_mm.c)."""
    Xa, ya, Xb, yb = [np.ascontiguousarray(a, dtype=np.float64) for a in (Xa, ya, Xb, yb)]
    quantiles = np.ascontiguousarray(quantiles, dtype=np.float64)
    K, na, nb, nq = Xa.shape[1], Xa.shape[0], Xb.shape[0], len(quantiles)
    S = 3 * nq
    taus = np.ascontiguousarray(taus, dtype=np.float64).reshape(reps + 1, sims)
    da = np.ascontiguousarray(draw_a, dtype=np.uint32).reshape(reps + 1, sims)
    db = np.ascontiguousarray(draw_b, dtype=np.uint32).reshape(reps + 1, sims)
    idx_a = np.ascontiguousarray(idx_a if reps else np.zeros((0, na)), dtype=np.uint32)
    idx_b = np.ascontiguousarray(idx_b if reps else np.zeros((0, nb)), dtype=np.uint32)
    out = dict(point_stats=np.full(S, np.nan), betas_a=np.empty((sims, K)), betas_b=np.empty((sims, K)),
               rep_stats=np.full((max(reps, 1), S), np.nan), rep_status=np.zeros(max(reps, 1), dtype=np.int32),
               se=np.empty(S), p=np.empty(S), ci_lo=np.empty(S), ci_hi=np.empty(S), t=np.empty(S))
    nok = C.c_int64()
    U32 = C.POINTER(C.c_uint32)
    L = lib()
    L.orc_mm_run.restype = C.c_int
    rc = L.orc_mm_run(C.c_int32(K), _dp(Xa), _dp(ya), C.c_int64(na), _dp(Xb), _dp(yb), C.c_int64(nb), C.c_int32(sims), C.c_int32(nq),
                      _dp(quantiles), C.c_int64(reps), idx_a.ctypes.data_as(U32), idx_b.ctypes.data_as(U32), _dp(taus),
                      da.ctypes.data_as(U32), db.ctypes.data_as(U32), C.c_int(nthreads), _dp(out["point_stats"]),
                      _dp(out["betas_a"]), _dp(out["betas_b"]), _dp(out["rep_stats"]),
                      out["rep_status"].ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nok), _dp(out["se"]), _dp(out["p"]),
                      _dp(out["ci_lo"]), _dp(out["ci_hi"]), _dp(out["t"]))
    if rc != 0:
        raise OracleError(rc)
    for k in ("rep_stats", "rep_status"):
        out[k] = out[k][:reps]
    out.update(n_ok=int(nok.value), S=S)
    return out
