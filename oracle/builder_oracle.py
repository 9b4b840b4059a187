"""Python restatement of the reference's frame handling in OaxacaBuilder::run (TEST INFRASTRUCTURE ONLY):
clean_dataframe (builder.rs:760-784), create_dummies_manual (:380-418), split_groups (:61-102), prepare_data
(:294-378) and the name-prefix Yun membership (normalization.rs:14-20, builder.rs:636-647).
Frames are dicts of lists (None = null)."""
from __future__ import annotations

import numpy as np


class BuilderError(Exception):
    def __init__(self, kind, msg):
        super().__init__(msg)
        self.kind = kind


def prepare(frame: dict, outcome, group, reference_group, predictors=(), categorical=(), normalize=(), weights=None):
    cols = [outcome, group] + list(predictors) + list(categorical) + ([weights] if weights else [])
    for c in cols:                                                  # builder.rs:774-778
        if c not in frame:
            raise BuilderError("ColumnNotFound", c)
    n0 = len(frame[outcome])
    keep = [i for i in range(n0) if all(frame[c][i] is not None and not (isinstance(frame[c][i], float) and np.isnan(frame[c][i]))
                                        for c in cols)]             # drop_nulls(cols) :780-782
    names = ["__ob_intercept__"] + list(predictors)
    dummies, counts, base = [], {}, {}
    for cat in categorical:                                          # :794-806 on the full cleaned frame
        vals = [frame[cat][i] for i in keep]
        levels = sorted(set(vals))                                   # :384-388
        counts[cat] = len(levels)
        base[cat] = f"{cat}_{levels[0]}"                             # :400
        for lv in levels[1:]:                                        # :402-409
            names.append(f"{cat}_{lv}")
            dummies.append(np.array([1.0 if v == lv else 0.0 for v in vals]))
    g = [frame[group][i] for i in keep]
    uniq = sorted(set(g))
    if len(uniq) < 2:                                                # :67-71
        raise BuilderError("InvalidGroupVariable", "Not enough groups for comparison")
    a_name = uniq[0] if uniq[0] != reference_group else uniq[1]      # :73-83
    code = np.array([0 if v == a_name else (1 if v == reference_group else 2) for v in g])
    X = np.stack([np.ones(len(keep))] + [np.array([frame[p][i] for i in keep], float) for p in predictors] + dummies, 1)
    y = np.array([frame[outcome][i] for i in keep], float)
    w = np.array([frame[weights][i] for i in keep], float) if weights else None
    A, B = code == 0, code == 1
    norm = []
    for var in normalize:                                            # normalization.rs:14-38
        idx = [i for i, nm in enumerate(names) if nm.startswith(var + "_")]
        norm.append(dict(m=counts.get(var, len(idx) + 1), idx=idx, has_base=var in base))
    return dict(names=names, base_names=[base[v] for v in normalize if v in base], n_cont=len(predictors),
                Xa=X[A], ya=y[A], wa=None if w is None else w[A], Xb=X[B], yb=y[B], wb=None if w is None else w[B],
                norm=norm, group=code, rows=len(keep))
