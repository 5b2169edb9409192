#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms, share.  (The per-launch times of
such a pass are cold-cache and serialised: the SHARES are what to compare with the bench line, not the absolute times.)
usage: launch_shares.py launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = collections.Counter(), collections.Counter()
for i, r in enumerate(csv.DictReader(rows)):
    if i < skip or r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v if unit == "ms" else v * 1e3
    tot[name] += ms
    cnt[name] += 1
all_ms = sum(tot.values()) or 1.0
print(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k:48s} {cnt[k]:8d} {v:10.3f} {100 * v / all_ms:6.1f}%")
print(f"{'all':48s} {sum(cnt.values()):8d} {all_ms:10.3f}")
