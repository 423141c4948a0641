"""profiles/ncu_traffic.json from the unit captures of tools/ncu_units.sh (second repetition of every unit):
per bench family, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged over the launches of
the family in the capture, with the capture named as the source.  bench.py copies it into `roofline.traffic`.
Usage: python tools/make_ncu_traffic.py gpurun_out/ncu_rhn.csv gpurun_out/ncu_cell_0.csv ... (files are also
summarised into profiles/r02_ncu_<unit>.csv)"""
import csv, json, os, re, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # family_of

fam_bytes, fam_n, fam_src = defaultdict(float), defaultdict(int), defaultdict(set)
for f in sys.argv[1:]:
    rows = [l for l in open(f, newline="") if l.startswith('"')]
    per, names = defaultdict(dict), {}
    for r in csv.DictReader(rows):
        i = int(r["ID"])
        names[i] = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r["Kernel Name"])).replace("gnas::", "")
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
            per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * mult
    ids = sorted(per)
    ids = ids[len(ids) // 2:]
    for i in ids:
        n = names[i]
        # ncu prints template arguments as <15, 1, 2>; bench's matcher expects the __PRETTY_FUNCTION__ form
        m = re.match(r"(?:\w+::)?(k_\w+)<(.*)>", n)
        key = n
        if m and m.group(1) in ("k_dense_fwd", "k_dense_bwd_dx"):
            kw = m.group(2).split(",")[0].strip()
            key = f"{m.group(1)} [with int KW = {kw}; ]"
        elif m and m.group(1) == "k_conv_tc":
            key = f"k_conv_tc [with ,{m.group(2).split(',')[1].strip()}>]"
        key = key.replace("rtc::", "").replace("tc::", "").replace("head::", "")
        fam, bound, _ = bench.family_of(key)
        if bound is None:
            continue
        fam_bytes[fam] += per[i].get("dram__bytes_read.sum", 0.0) + per[i].get("dram__bytes_write.sum", 0.0)
        fam_n[fam] += 1
        fam_src[fam].add("profiles/r02_" + os.path.basename(f))
out = {f: {"bytes_per_launch": round(fam_bytes[f] / fam_n[f]), "launches_in_capture": fam_n[f],
           "source": ", ".join(sorted(fam_src[f])) + " (stand-alone unit, cold L2: RHN layer T=42 B=64 H=512 / search cells 0, 3, 5 at B=64)"}
       for f in sorted(fam_bytes)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
