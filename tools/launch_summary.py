"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv <cmd>`) per kernel:
count, total and average duration, share of the summed GPU time.  No GPU needed.
usage: python tools/launch_summary.py gpurun_out/launches.csv [title] > profiles/launches_summary.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(sys.argv[1])) if r and r[0] == "ID")
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = r[ik][:110]
    tot[name] += v
    cnt[name] += 1
allt = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]} (gpu__time_duration.sum; cold-cache, serialised: compare shares)\n")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{name:110s} n={cnt[name]:4d} total={t:10.1f} us avg={t / cnt[name]:9.1f} us share={100 * t / allt:5.1f}%")
