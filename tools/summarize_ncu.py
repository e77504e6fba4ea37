#!/usr/bin/env python
"""Per-kernel HBM roofline table from an `ncu --page raw --csv` export (the files profiles/*_raw.csv).

    python tools/summarize_ncu.py profiles/r02_ncu_elementwise_i_raw.csv [more.csv ...] [--peak 6480.8]

For every kernel name: launches in the capture, summed duration, summed DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), achieved DRAM GB/s and its fraction of the measured copy rate (MEASURED_PEAKS.json hbm_gbs).
ncu serialises the launches and replays each one with cold caches, so these are per-kernel figures, not step shares.
"""
import argparse
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def num(v: str) -> float:
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--peak", type=float, default=None, help="HBM GB/s (default: MEASURED_PEAKS.json hbm_gbs)")
    a = ap.parse_args()
    peak = a.peak
    if peak is None:
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except (OSError, KeyError, ValueError):
            peak = 6480.8
    for path in a.files:
        rows = list(csv.reader(open(path)))
        head, units = rows[0], rows[1]
        col = {n: head.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
        agg = OrderedDict()
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").split("(")[0]
            ms = num(r[col["gpu__time_duration.sum"]]) * TO_MS[units[col["gpu__time_duration.sum"]]]
            by = sum(num(r[col[k]]) * TO_BYTES[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            e = agg.setdefault(name, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += ms
            e[2] += by
        print(f"# {os.path.relpath(path, ROOT) if path.startswith(ROOT) else path}   (HBM copy rate {peak:.0f} GB/s)")
        print(f"{'kernel':44s} {'launches':>8s} {'ms':>9s} {'DRAM MB':>10s} {'GB/s':>8s} {'of copy rate':>12s}")
        for name, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = by / (ms * 1e-3) / 1e9 if ms else 0.0
            print(f"{name[:44]:44s} {n:8d} {ms:9.3f} {by / 1e6:10.1f} {gbs:8.0f} {gbs / peak:12.2f}")
        print()


if __name__ == "__main__":
    sys.exit(main())
