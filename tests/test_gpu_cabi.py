"""The boundary from plain C: tests/c/abi_smoke.c is compiled with gcc against include/obboot.h and libobboot.so (no
Python, no torch in that process), run on the GPU, and its numbers must equal what the ctypes binding gets for the same
frame (restated here in numpy from the C program's LCG) bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "oaxaca_blinder_rs_b200", "_lib")


def build_c(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", LIBDIR, "-lobboot", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", exe])
    return exe


def test_c_program_compiles_and_links_against_the_header(tmp_path):
    """CPU: the C translation unit builds against the public header alone (C, not C++) and links the library."""
    from oaxaca_blinder_rs_b200 import _native
    _native.build()
    exe = build_c(tmp_path)
    r = subprocess.run([exe, "100", "4", "1"], capture_output=True, text=True)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def lcg_frame(n):
    state = np.uint64(0x0B200)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    u = np.empty(6 * n)
    with np.errstate(over="ignore"):
        for i in range(6 * n):
            state = state * a + c
            u[i] = float(state >> np.uint64(11)) / 9007199254740992.0
    u = u.reshape(n, 6)
    grp = np.where(u[:, 0] < 0.5, 0, 1).astype(np.uint8)
    x0 = 8.0 + 12.0 * u[:, 1] + np.where(grp == 0, 0.5, 0.0)
    x1 = 40.0 * u[:, 2]
    cat = np.where(u[:, 3] < 0.4, 0, np.where(u[:, 3] < 0.75, 1, 2)).astype(np.int32)
    w = 0.5 + 2.5 * u[:, 4]
    y = np.where(grp == 0, 2.9, 2.7) + 0.08 * x0 + 0.01 * x1 + 0.1 * cat + (u[:, 5] - 0.5)
    return grp, x0, x1, cat, w, y


@pytest.mark.gpu
def test_c_program_matches_the_ctypes_binding(tmp_path):
    import oaxaca_blinder_rs_b200 as ob
    n, reps, seed = 20_000, 64, 7
    exe = build_c(tmp_path)
    r = subprocess.run([exe, str(n), str(reps), str(seed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    vals = {}
    for line in r.stdout.split("\n"):
        if line:
            k, v = line.split()
            vals.setdefault(k, []).append(float(v))
    assert vals["async_equals_sync"] == [1.0]
    grp, x0, x1, cat, w, y = lcg_frame(n)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, [x0, x1], [cat], [3], y, w, grp)
    out = ob.bootstrap(des, reps, ref_kind=ob.REF_POOLED, norm=[ob.NormVar(3, [3, 4])], seed=seed)
    des.close(); ctx.close()
    assert (vals["n_a"][0], vals["n_b"][0], vals["n_ok"][0]) == (des.n_a, des.n_b, out["n_ok"])
    assert vals["total_gap"][0] == out["total_gap"]
    assert np.array_equal(vals["point"], out["point_stats"]) and np.array_equal(vals["se"], out["std_err"])
    assert np.array_equal(vals["ci_lo"], out["ci_lower"]) and np.array_equal(vals["ci_hi"], out["ci_upper"])
    assert np.array_equal(vals["beta_star"], out["beta_star"])
    rs = 0.0
    for v in np.abs(out["residuals_b"]).tolist():          # the C program's left-to-right sum
        rs += v
    assert vals["resid_abs_sum"][0] == rs
    # ob_mm_run from C == the ctypes binding, bit for bit (native streams keyed by the seed)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, [x0, x1], [cat], [3], y, None, grp)
    mm = ob.machado_mata(des, [0.25, 0.5, 0.75], simulations=24, reps=5, seed=seed)
    des.close(); ctx.close()
    assert (vals["mm_n_ok"][0], vals["mm_qr_total"][0], vals["mm_qr_failed"][0]) == (mm["n_ok"], 2 * 6 * 24, 0)
    assert np.array_equal(vals["mm_point"], mm["point_stats"].ravel()) and np.array_equal(vals["mm_se"], mm["std_err"].ravel())
    assert np.array_equal(vals["mm_ci_lo"], mm["ci_lower"].ravel()) and np.array_equal(vals["mm_ci_hi"], mm["ci_upper"].ravel())
