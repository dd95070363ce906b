"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list of
tests/prof_step.py: per kernel name the launches, time and DRAM bytes of the LAST training step in the capture, the
share of the step, and the mean DRAM traffic per launch of the igemm family (-> profiles/roofline_traffic.json).
usage: python profiles/summarize_launches.py launches.csv [--write-traffic]"""
import csv
import json
import os
import re
import sys

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
launch = {}
for r in rows:
    d = launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1e-3)
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
ids = sorted(launch)
# the last step = from the last conv1_fwd_kernel launch to the end
starts = [i for i in ids if "conv1_fwd_kernel" in launch[i]["name"]]
first = starts[-1]
step = [launch[i] for i in ids if i >= first]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("ub::", "")


agg = {}
for d in step:
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
print(f"last step: {len(step)} launches, {total / 1e3:.3f} ms (ncu per-launch times: cold caches, serialised -- compare shares)")
print(f"{'kernel':60s} {'n':>4s} {'us':>10s} {'share':>7s} {'DRAM MB':>10s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {a[0]:4d} {a[1]:10.1f} {100 * a[1] / total:6.1f}% {a[2] / 1e6:10.1f}")
ig = [(a[0], a[2]) for k, a in agg.items() if k.startswith("igemm")]
n_ig, by_ig = sum(x[0] for x in ig), sum(x[1] for x in ig)
if n_ig and by_ig:
    print(f"igemm family: {n_ig} launches, {by_ig / n_ig / 1e6:.1f} MB DRAM per launch")
    if "--write-traffic" in sys.argv:
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roofline_traffic.json")
        with open(out, "w") as f:
            json.dump({"igemm_dram_bytes_per_launch": by_ig / n_ig, "igemm_launches_per_step": n_ig,
                       "source": f"profiles/{os.path.basename(path)} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over "
                                 "the igemm-family launches of one batch-16 training step)"}, f)
