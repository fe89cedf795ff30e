#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/launch_shares.py gpurun_out/launches_TAG.csv [--ours]"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    if name.startswith("k_bench_"):      # apb_bench_peaks microbenchmarks, not part of the step
        continue
    if "--ours" in sys.argv and not name.startswith("k_"):
        name = "(torch / other)"
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {k[:60]} | {cnt[k]} | {v:.3f} | {v / s:.3f} | {1e3 * v / cnt[k]:.1f} |")
