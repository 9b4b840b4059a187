"""Host-side OaxacaBuilder mirror (C++ csrc/host/builder.cc through include/obboot_builder.h), CPU only:
cleaning, dummy coding, group rule, names, .normalize membership, formula parsing, error variants --
against the Python restatement of builder.rs in oracle/builder_oracle.py and the reference's own tests."""
import numpy as np
import pytest

import oaxaca_blinder_rs_b200 as ob
from oracle import builder_oracle as bo


def frame_with_everything():
    rng = np.random.default_rng(3)
    n = 60
    return {
        "wage": [None if i == 7 else float(v) for i, v in enumerate(rng.normal(20, 3, n))],
        "education": [None if i == 11 else float(v) for i, v in enumerate(rng.normal(13, 2, n))],
        "experience": list(map(float, rng.uniform(0, 40, n))),
        "gender": [None if i == 13 else ("X" if i % 19 == 5 else ("M", "F")[int(v)]) for i, v in enumerate(rng.integers(0, 2, n))],
        "sector": [("tech", "agri", "serv", "manu")[i % 4] for i in range(n)],
        "sector_size": list(map(float, rng.uniform(1, 9, n))),     # swept in by the "sector_" prefix
        "region": [("north", "south")[(i // 3) % 2] for i in range(n)],
        "w": list(map(float, rng.uniform(0.5, 3.0, n))),
    }


def test_describe_matches_builder_rules():
    f = frame_with_everything()
    b = ob.OaxacaBuilder(f, "wage", "gender", "F")
    b.predictors(["education", "experience", "sector_size"]).categorical_predictors(["sector", "region"]) \
        .normalize(["sector", "region", "experience"]).weights("w")
    d = b.describe()
    e = bo.prepare(f, "wage", "gender", "F", ["education", "experience", "sector_size"], ["sector", "region"],
                   ["sector", "region", "experience"], "w")
    assert d["names"] == e["names"] == ["__ob_intercept__", "education", "experience", "sector_size", "sector_manu",
                                        "sector_serv", "sector_tech", "region_south"]
    assert d["base_names"] == e["base_names"] == ["sector_agri", "region_north"]
    assert d["rows"] == e["rows"] == 57 and d["group"] == list(e["group"])
    assert (d["n_a"], d["n_b"]) == (len(e["ya"]), len(e["yb"]))
    assert d["cat_levels"] == [4, 2]
    # name-prefix membership: "sector_" also matches the continuous predictor sector_size (normalization.rs:14-20);
    # "experience" is not categorical: no dummies -> m = 0 + 1, no base row (builder.rs:636-640)
    off = d["norm_off"]
    got = [dict(m=d["norm_m"][v], idx=d["norm_idx"][off[v]:off[v + 1]], has_base=bool(d["norm_has_base"][v])) for v in range(3)]
    assert got == e["norm"]
    assert got[0]["idx"] == [3, 4, 5, 6] and got[2] == dict(m=1, idx=[], has_base=False)


def test_group_a_is_first_sorted_non_reference_value():
    f = {"y": [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], "g": ["b", "c", "a", "a", "b", "c"], "x": [1.0, 2.0, 3.0, 4.0, 5.0, 7.0]}
    d = ob.OaxacaBuilder(f, "y", "g", "a").predictors(["x"]).describe()
    assert d["group"] == [0, 2, 1, 1, 0, 2]           # A = "b" (first sorted value that is not the reference), "c" ignored
    d = ob.OaxacaBuilder(f, "y", "g", "c").predictors(["x"]).describe()
    assert d["group"] == [2, 1, 0, 0, 2, 1]           # A = "a"


def test_null_handling_counts(golden):
    k = golden["KAT"]["null_handling"]                 # null_handling_test.rs:4-64
    f = {"outcome": k["outcome"], "group": k["group"], "education": k["education"]}
    d = ob.OaxacaBuilder(f, "outcome", "group", "B").predictors(["education"]).describe()
    assert (d["n_a"], d["n_b"]) == (k["n_a"], k["n_b"]) == (3, 3)


def test_error_variants():
    f = {"y": [1.0, 2.0], "g": ["a", "b"], "x": [1.0, 2.0], "s": ["u", "v"]}
    with pytest.raises(ob.OaxacaError) as e:           # builder.rs:776
        ob.OaxacaBuilder(f, "y", "g", "a").predictors(["missing"]).describe()
    assert e.value.kind == "ColumnNotFound" and str(e.value) == "Column not found: missing"
    with pytest.raises(ob.OaxacaError) as e:           # builder.rs:67-71
        ob.OaxacaBuilder({"y": [1.0, 2.0], "g": ["a", "a"], "x": [1.0, 2.0]}, "y", "g", "a").predictors(["x"]).describe()
    assert e.value.kind == "InvalidGroupVariable" and "Not enough groups" in str(e.value)
    with pytest.raises(ob.OaxacaError) as e:           # outcome must be Float64 (builder.rs:308)
        ob.OaxacaBuilder(f, "s", "g", "a").predictors(["x"]).describe()
    assert e.value.kind == "PolarsError"
    with pytest.raises(ob.OaxacaError) as e:           # group must be String (builder.rs:75)
        ob.OaxacaBuilder(f, "y", "x", "a").predictors(["x"]).describe()
    assert e.value.kind == "PolarsError"


@pytest.mark.parametrize("formula,exp", [
    ("wage ~ education + experience + C(sector)", ("wage", ["education", "experience"], ["sector"])),
    ("y~ a+factor( b )+C(c)", ("y", ["a"], ["b", "c"])),
])
def test_formula(formula, exp):                        # formula.rs:12-60, formula_test.rs:4-27
    cols = {n: [1.0, 2.0] for n in ["wage", "education", "experience", "y", "a"]}
    cols.update({n: ["p", "q"] for n in ["sector", "b", "c", "g"]})
    d = ob.OaxacaBuilder.from_formula(cols, formula, "g", "p").describe()
    outcome, preds, cats = exp
    assert d["names"][1:1 + len(preds)] == preds and len(d["cat_levels"]) == len(cats)


@pytest.mark.parametrize("bad", ["wage education", "~ x", "y ~ ", "a ~ b ~ c"])
def test_formula_errors(bad):
    with pytest.raises(ob.OaxacaError) as e:
        ob.OaxacaBuilder.from_formula({"y": [1.0]}, bad, "g", "p")
    assert e.value.kind == "InvalidGroupVariable"


def test_builder_symbols_exported():
    import os, re
    from oaxaca_blinder_rs_b200 import _native, builder
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "obboot_builder.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ob_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(builder.BUILDER_SYMBOLS)
    for s in declared:
        assert hasattr(_native.lib(), s), s


def test_csv_reader_matches_python_restatement(tmp_path):
    """DataFrame::read_csv (single pass, numeric columns parsed straight to f64) against a restatement with the csv module:
    quotes and escaped quotes, CRLF, blank lines, short rows, empty cells = nulls, '+1' / ' 2' / 'inf' accepted like strtod,
    a column that turns out non-numeric late, an all-empty column."""
    import csv
    import io
    import oaxaca_blinder_rs_b200 as ob
    rng = np.random.default_rng(0)
    n = 400
    lines = ["wage,edu,exp,sector,gender,late,empty,\"quoted,name\""]
    for i in range(n):
        wage = f"{rng.normal(10, 2):.6f}"
        edu = ["+12", " 13", "1.4e1", "15", ""][i % 5]                  # strtod-compatible spellings, a null
        exp = f"{rng.uniform(0, 40):.3f}" if i % 37 else ""
        sector = ['agri', '"ma""nu"', '"serv,ices"', 'tech'][i % 4]
        gender = "M" if rng.random() < 0.5 else "F"
        late = "7" if i < 300 else "seven"                              # numeric for 300 rows, then text
        row = f"{wage},{edu},{exp},{sector},{gender},{late},,{i}"
        if i % 50 == 49:
            row = f"{wage},{edu},{exp},{sector},{gender}"                # short row: missing cells are nulls
        lines.append(row + ("\r" if i % 3 == 0 else ""))
        if i % 97 == 0:
            lines.append("")                                             # blank line
    path = tmp_path / "tricky.csv"
    path.write_text("\n".join(lines) + "\n")
    # python restatement
    rows = [r for r in csv.reader(io.StringIO(path.read_text().replace("\r", ""))) if r]
    hdr, rows = rows[0], [r + [""] * (8 - len(r)) for r in rows[1:]]
    col = {h: [r[j] for r in rows] for j, h in enumerate(hdr)}
    keep = [i for i in range(len(rows)) if all(col[c][i] != "" for c in ("wage", "gender", "edu", "exp", "sector", "late"))]
    frame = ob.read_csv(str(path))
    b = ob.OaxacaBuilder(frame, "wage", "gender", "F").predictors(["edu", "exp"]).categorical_predictors(["sector", "late"])
    d = b.describe()
    assert d["rows"] == len(keep)
    assert d["names"] == ["__ob_intercept__", "edu", "exp", 'sector_ma"nu', "sector_serv,ices", "sector_tech", "late_seven"]
    assert d["base_names"] == [] and d["cat_levels"] == [4, 2]
    assert abs(d["outcome_sum"] - sum(float(col["wage"][i]) for i in keep)) < 1e-9
    assert abs(d["cont_sums"][0] - sum(float(col["edu"][i]) for i in keep)) < 1e-9
    assert abs(d["cont_sums"][1] - sum(float(col["exp"][i]) for i in keep)) < 1e-9
    assert d["group"] == [0 if col["gender"][i] == "M" else 1 for i in keep]
    # the all-empty column and the quoted header exist; using the empty one as a predictor drops every row
    e = ob.OaxacaBuilder(frame, "wage", "gender", "F").predictors(["quoted,name"]).describe()
    assert e["rows"] == len([i for i in range(len(rows)) if col["quoted,name"][i] != "" and col["wage"][i] != ""])
    with pytest.raises(ob.OaxacaError):
        ob.read_csv(str(tmp_path / "missing.csv"))


def test_pandas_missing_values_are_nulls_not_levels():
    """pd.NA ('string' / nullable dtypes), NaN and None in string columns are nulls: the row is dropped
    (clean_dataframe, builder.rs:760-784), never encoded as a level called "<NA>"."""
    pd = pytest.importorskip("pandas")
    n = 12
    base = {"wage": [float(10 + i) for i in range(n)], "education": [float(8 + (i * 7) % 5) for i in range(n)],
            "gender": ["M", "F"] * (n // 2), "sector": ["agri", "tech", "serv"] * (n // 3)}
    ref = ob.OaxacaBuilder(dict(base, gender=[None if i == 2 else g for i, g in enumerate(base["gender"])],
                                sector=[None if i == 5 else s for i, s in enumerate(base["sector"])]),
                           "wage", "gender", "F").predictors(["education"]).categorical_predictors(["sector"]).describe()
    for dtype in ("string", "object", "category"):
        df = pd.DataFrame(base)
        df["gender"] = df["gender"].astype(dtype)
        df["sector"] = df["sector"].astype(dtype)
        df.loc[2, "gender"] = pd.NA if dtype == "string" else None
        df.loc[5, "sector"] = pd.NA if dtype == "string" else np.nan
        df["education"] = df["education"].astype("Float64")        # nullable numeric, no nulls here
        d = ob.OaxacaBuilder(df, "wage", "gender", "F").predictors(["education"]).categorical_predictors(["sector"]).describe()
        assert d["rows"] == n - 2 and d["cat_levels"] == [3], dtype
        assert d["names"] == ref["names"] and d["group"] == ref["group"] and (d["n_a"], d["n_b"]) == (ref["n_a"], ref["n_b"]), dtype
        assert not any("<NA>" in nm or "nan" in nm.lower() for nm in d["names"] + d["base_names"]), dtype


@pytest.mark.parametrize("flag,value", [("--ref-coeffs", "cotton"), ("--ref-coeffs", "Pooled"), ("--bootstrap-reps", "abc"),
                                        ("--bootstrap-reps", "-5"), ("--rif-quantile", "1.5"), ("--rif-quantile", "x"),
                                        ("--seed", "1e3")])
def test_cli_rejects_untyped_values(flag, value):
    """clap's ValueEnum / typed args (main.rs:44-128) reject these with exit code 2; a typo must never fall back
    silently to another reference-coefficient kind or to zero replicates."""
    import os
    import subprocess
    from oaxaca_blinder_rs_b200 import _native
    _native.build()
    cli = os.path.join(os.path.dirname(_native.LIB_PATH), "oaxaca-cli")
    r = subprocess.run([cli, "--data", "nope.csv", "--outcome", "y", "--group", "g", "--reference", "r", flag, value],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "invalid value" in r.stderr and value in r.stderr


# ---- Machado-Mata builder mirror: host logic that needs no device ----
def test_quantile_decomposition_key_format():
    """format!("q{}", (tau * 100.0) as u32), quantile_decomposition.rs:277: truncation, not rounding (0.29 * 100 =
    28.999999999999996 -> "q28"), saturation of negative values / NaN to 0 like Rust's float -> u32 cast."""
    import oaxaca_blinder_rs_b200 as ob
    key = ob.QuantileDecompositionBuilder.quantile_key
    assert [key(t) for t in (0.1, 0.25, 0.5, 0.75, 0.9)] == ["q10", "q25", "q50", "q75", "q90"]     # the defaults (:57)
    assert key(0.29) == "q28" and key(0.57) == "q56" and key(0.58) == "q57"
    assert key(0.999) == "q99" and key(1.0) == "q100" and key(0.0) == "q0"
    assert key(-0.3) == "q0" and key(float("nan")) == "q0"
    for t in (0.07, 0.14, 0.28, 0.55, 0.56):
        assert key(t) == "q%d" % int(t * 100.0)


def test_quantile_decomposition_null_handling():
    """The Machado-Mata builder does not clean the frame (quantile_decomposition.rs:286-287 only selects columns): a null
    outcome in one of the two groups is InvalidGroupVariable("Null outcome encountered") (prepare_data, :111-118), a null
    predictor fails the ndarray conversion (:141).  Both are found before any device work."""
    import oaxaca_blinder_rs_b200 as ob
    base = {"wage": [10.0, 12.0, 11.0, 13.0, 20.0, 22.0, 21.0, 23.0], "education": [12.0, 16.0, 14.0, 16.0, 12.0, 16.0, 14.0, 16.0],
            "gender": ["F", "F", "F", "F", "M", "M", "M", "M"]}
    f = dict(base); f["wage"] = list(base["wage"]); f["wage"][2] = None
    with pytest.raises(ob.OaxacaError) as e:
        ob.QuantileDecompositionBuilder(f, "wage", "gender", "F").predictors(["education"]).simulations(10).bootstrap_reps(1).run()
    assert e.value.kind == "InvalidGroupVariable" and "Null outcome encountered" in str(e.value)
    f = dict(base); f["education"] = list(base["education"]); f["education"][5] = None
    with pytest.raises(ob.OaxacaError) as e:
        ob.QuantileDecompositionBuilder(f, "wage", "gender", "F").predictors(["education"]).simulations(10).bootstrap_reps(1).run()
    assert e.value.kind == "PolarsError"
    # a null in a row of a THIRD group never reaches a design (:195-206): no null error; the run then needs the device
    f = dict(base)
    f["gender"] = base["gender"] + ["X"]; f["wage"] = base["wage"] + [None]; f["education"] = base["education"] + [12.0]
    import torch
    b = ob.QuantileDecompositionBuilder(f, "wage", "gender", "F").predictors(["education"]).simulations(10).bootstrap_reps(1)
    if torch.cuda.is_available():
        assert (b.run().n_a, ) == (4, )
    else:
        with pytest.raises(ob.OaxacaError) as e:
            b.run()
        assert e.value.kind == "NoDevice"
