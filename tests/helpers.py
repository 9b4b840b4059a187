"""Shared test helpers: host-side restatement of the frame -> design step for the small fixtures."""
import numpy as np


def fixture_design(fix, weighted=False):
    c = fix["columns"]
    g = np.array(c[fix["group"]])
    refg = fix["reference_group"]
    other = sorted(set(g) - {refg})[0]          # builder.rs:73-83
    A, B = g == other, g == refg
    cols = [np.ones(len(g))] + [np.array(c[p], float) for p in fix["predictors"]]
    norm = []
    for cat in fix.get("categorical", []):
        v = np.array(c[cat])
        levels = sorted(set(v))                   # builder.rs:384-388
        start = len(cols)
        for lv in levels[1:]:                     # builder.rs:402
            cols.append((v == lv).astype(float))
        if cat in fix.get("normalize", []):
            norm.append((len(levels), list(range(start, len(cols)))))
    X = np.stack(cols, 1)
    y = np.array(c[fix["outcome"]], float)
    w = np.array(c["weight"], float) if weighted else None
    return (X[A], y[A], None if w is None else w[A], X[B], y[B], None if w is None else w[B],
            norm, len(fix["predictors"]))


