"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total time, share.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [header comment] > profiles/rNN_launches_summary.csv"""
import csv, re, sys
from collections import defaultdict
rows = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
rd = csv.DictReader(rows)
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name).replace("gnas::", "")
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000.0 if r["Metric Unit"] == "ns" else (v * 1000.0 if r["Metric Unit"] == "ms" else v)   # -> us
    tot[name] += v
    cnt[name] += 1
allt = sum(tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,total_us,share_pct,avg_us")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"\"{k}\",{cnt[k]},{tot[k]:.1f},{100 * tot[k] / allt:.2f},{tot[k] / cnt[k]:.2f}")
