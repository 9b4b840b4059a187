#!/usr/bin/env python
"""Summarises an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>):
per kernel the number of launches, the summed duration and its share.  Usage: launch_summary.py X.csv "<command>" > summary.json"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    acc = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*$", "", r[ci["Kernel Name"]]).replace("ob::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        acc[name][0] += 1
        acc[name][1] += ms
    total = sum(v[1] for v in acc.values())
    kernels = [dict(kernel=k, launches=v[0], ms=round(v[1], 3), share=round(v[1] / total, 5)) for k, v in
               sorted(acc.items(), key=lambda kv: -kv[1][1])]
    json.dump(dict(command=cmd, total_ms=round(total, 3), kernels=kernels), sys.stdout, indent=1)


if __name__ == "__main__":
    main()
