"""Per-kernel summary of a tools/ncu_units.sh capture (long-format ncu CSV): launches, time, DRAM / L2 bytes,
issue activity, main stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/ncu_cell_0.csv [skip_first_half]"""
import csv, re, sys
from collections import defaultdict
rows = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(rows):
    i = int(r["ID"])
    names[i] = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r["Kernel Name"])).replace("gnas::", "")
    try:
        per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[i]["unit:" + r["Metric Name"]] = r["Metric Unit"]
    except ValueError:
        pass
ids = sorted(per)
if len(sys.argv) > 2:
    ids = ids[len(ids) // 2:]          # second repetition only
agg = defaultdict(lambda: defaultdict(float))
for i in ids:
    m, a = per[i], agg[names[i]]
    t = m.get("gpu__time_duration.sum", 0.0)
    u = m.get("unit:gpu__time_duration.sum", "ns")
    t = t / 1000.0 if u in ("ns", "nsecond") else (t * 1000 if u.startswith("ms") else t)
    a["n"] += 1; a["us"] += t
    for k, sc in (("dram__bytes_read.sum", 1), ("dram__bytes_write.sum", 1), ("lts__t_bytes.sum", 1)):
        v, un = m.get(k, 0.0), m.get("unit:" + k, "byte")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
        a[k] += v * mult
    for k in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__grid_size",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"):
        a[k] += m.get(k, 0.0) * t
print("kernel,launches,total_us,avg_us,dram_MB_per_launch,l2_MB_per_launch,dram_GBps,issue_active_pct,warps_active_pct,regs,grid,stall_long_sb,stall_short_sb,stall_lg,stall_mio,stall_barrier")
for k in sorted(agg, key=lambda k: -agg[k]["us"]):
    a = agg[k]; n, us = a["n"], max(a["us"], 1e-9)
    dram = a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]
    w = lambda key: a[key] / us
    print(f"\"{k}\",{int(n)},{us:.1f},{us / n:.2f},{dram / n / 1e6:.2f},{a['lts__t_bytes.sum'] / n / 1e6:.2f},{dram / us / 1e3:.0f},"
          f"{w('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f},{w('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f},"
          f"{w('launch__registers_per_thread'):.0f},{w('launch__grid_size'):.0f},"
          f"{w('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.2f},"
          f"{w('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'):.2f},"
          f"{w('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'):.2f},"
          f"{w('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'):.2f},"
          f"{w('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):.2f}")
