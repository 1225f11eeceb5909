"""ncu launch list of one iteration (scripts/one_iter.py) -> per-kernel-family table + profiles/traffic.json.

    python scripts/summarize_iter.py gpurun_out/r2_iter_launches.csv profiles/r2_iter_launches_summary.txt
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

src, out_txt = sys.argv[1], sys.argv[2]
rows = []
with open(src) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    rows.append(r)
fam = defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
launch = {}


def unit_scale(unit, to):
    t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6} if to == "us" else \
        {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return t.get(unit, 1.0)


for r in rows:
    key = r["ID"]
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("acg::", "").replace("(anonymous namespace)::", "")
    d = launch.setdefault(key, {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * unit_scale(r["Metric Unit"], "us")
    elif r["Metric Name"] == "dram__bytes_read.sum":
        d["rd"] = v * unit_scale(r["Metric Unit"], "B")
    elif r["Metric Name"] == "dram__bytes_write.sum":
        d["wr"] = v * unit_scale(r["Metric Unit"], "B")
for d in launch.values():
    f = fam[d["name"]]
    f["n"] += 1
    f["us"] += d["us"]
    f["rd"] += d["rd"]
    f["wr"] += d["wr"]
tot_us = sum(f["us"] for f in fam.values())
with open(out_txt, "w") as fh:
    fh.write("%-62s %5s %10s %7s %10s %10s\n" % ("kernel", "n", "time us", "share", "DRAM rd MB", "DRAM wr MB"))
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        fh.write("%-62s %5d %10.1f %6.1f%% %10.1f %10.1f\n" % (name[:62], f["n"], f["us"], 100 * f["us"] / tot_us,
                                                              f["rd"] / 1e6, f["wr"] / 1e6))
    fh.write("%-62s %5d %10.1f\n" % ("total", sum(f["n"] for f in fam.values()), tot_us))
conv = {k: v for k, v in fam.items() if "conv_" in k and "pack" not in k}
conv_bytes = sum(v["rd"] + v["wr"] for v in conv.values())
all_bytes = sum(v["rd"] + v["wr"] for v in fam.values())
js = {"conv_dram_bytes_per_iteration": conv_bytes, "all_kernels_dram_bytes_per_iteration": all_bytes,
      "conv_kernel_time_us_cold": sum(v["us"] for v in conv.values()), "launches": len(launch),
      "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the tcgen05 conv kernels of ONE iteration "
              "(train_d + train_g, B=256, eager, one stream, cold cache per launch); source " + os.path.basename(src)}
json.dump(js, open(os.path.join(os.path.dirname(out_txt), "traffic.json"), "w"), indent=1)
print(open(out_txt).read())
print(json.dumps(js))
