#!/usr/bin/env python
"""Turns an ncu csv of the Gram kernel's DRAM counters into an entry of profiles/gram_traffic.json (the file
bench.py's roofline.traffic reads).  The entry is keyed by the sha256 of the kernel sources, so it stops counting the
moment gram.cu / internal.h / common.cuh change.

  on the GPU box (one launch of the default bench command, after the same command exited 0 without ncu):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gram_ws_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_gram_dram_config3.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline
  here:
    python tools/gram_traffic.py gpurun_out/r02_gram_dram_config3.csv config3_n10M_k50_wls_yun_B2000 1
"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    head = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    cols = {c: i for i, c in enumerate(rows[head])}
    out = {}
    for r in rows[head + 1:]:
        name, unit, val = r[cols["Metric Name"]], r[cols["Metric Unit"]], float(r[cols["Metric Value"]].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6,
                 "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        out.setdefault(name, []).append(val * scale)
        out["kernel"] = r[cols["Kernel Name"]]
    return out


def main():
    src, workload, n_gpus = sys.argv[1], sys.argv[2], int(sys.argv[3])
    import bench
    m = parse(src)
    total = m["dram__bytes_read.sum"][0] + m["dram__bytes_write.sum"][0]
    dst_name = "r02_" + os.path.basename(src).replace("r02_", "")
    shutil.copy(src, os.path.join(ROOT, "profiles", dst_name))
    path = os.path.join(ROOT, "profiles", "gram_traffic.json")
    try:
        doc = json.load(open(path))
    except (OSError, ValueError):
        doc = {"what": "ncu dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the Gram kernel, per workload; "
                       "an entry counts only while src_sha256 equals the hash of the current kernel sources", "entries": []}
    doc["entries"] = [e for e in doc["entries"] if not (e["workload"] == workload and e["n_gpus"] == n_gpus)]
    doc["entries"].append({"workload": workload, "n_gpus": n_gpus, "bytes": int(total),
                           "read_bytes": int(m["dram__bytes_read.sum"][0]), "write_bytes": int(m["dram__bytes_write.sum"][0]),
                           "kernel": m.get("kernel"), "ncu_duration_ms": m.get("gpu__time_duration.sum", [0])[0] / 1e6,
                           "src_sha256": bench.gram_source_hash(), "csv": "profiles/" + dst_name})
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(doc["entries"][-1], indent=1))


if __name__ == "__main__":
    main()
