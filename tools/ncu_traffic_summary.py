"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel: launches, total time, DRAM bytes (per iteration and per launch).  usage: ncu_traffic_summary.py launches.csv iterations out.json"""
import csv, json, re, sys
from collections import defaultdict
path, iters, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = defaultdict(lambda: dict(launches=0, ns=0.0, rd=0.0, wr=0.0))
unit_scale = {"nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ids = set()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1].strip()
    name = re.sub(r"<.*", "", name)
    m, u, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    a = acc[name]
    if m == "gpu__time_duration.sum":
        a["ns"] += v * unit_scale.get(u, 1.0)
        a["launches"] += 1
    elif m == "dram__bytes_read.sum":
        a["rd"] += v * unit_scale.get(u, 1.0)
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * unit_scale.get(u, 1.0)
res = {}
for k, a in sorted(acc.items(), key=lambda kv: -kv[1]["ns"]):
    if a["launches"] == 0:
        continue
    res[k] = dict(launches_per_iteration=a["launches"] / iters, ms_per_iteration_under_ncu=a["ns"] / 1e6 / iters,
                  dram_read_gb_per_iteration=a["rd"] / 1e9 / iters, dram_write_gb_per_iteration=a["wr"] / 1e9 / iters,
                  dram_bytes_per_launch=(a["rd"] + a["wr"]) / a["launches"],
                  dram_gbs_while_running=(a["rd"] + a["wr"]) / a["ns"] if a["ns"] else None)
json.dump(dict(source=" ".join(sys.argv), iterations_in_capture=iters, kernels=res), open(out, "w"), indent=1)
for k, v in res.items():
    print(f"{k:32s} n/iter={v['launches_per_iteration']:8.1f} ms/iter={v['ms_per_iteration_under_ncu']:9.3f} rd={v['dram_read_gb_per_iteration']:8.2f} GB wr={v['dram_write_gb_per_iteration']:8.2f} GB")
