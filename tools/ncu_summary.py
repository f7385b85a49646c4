# -*- coding: utf-8 -*-
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python tools/ncu_summary.py gpurun_out/launches.csv > profiles/launches_summary.txt"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("# source: {0}   (per-launch times under ncu are cold-cache and serialised: compare SHARES)".format(sys.argv[1]))
print("# total {0:.3f} ms over {1} launches".format(T / 1e6, sum(cnt.values())))
print("{0:40s} {1:>8s} {2:>12s} {3:>8s} {4:>12s}".format("kernel", "launches", "total_ms", "share%", "avg_us"))
for k, v in tot.most_common():
    print("{0:40s} {1:8d} {2:12.3f} {3:8.2f} {4:12.1f}".format(k, cnt[k], v / 1e6, 100 * v / T, v / cnt[k] / 1e3))
