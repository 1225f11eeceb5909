"""`ncu --set full` report of scripts/ncu_kernels.py -> one line per profiled launch with the metrics DESIGN.md quotes.

    ncu -i gpurun_out/r2_kernels.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/summarize_ncu_full.py /tmp/raw.csv gpurun_out/r2x_kernels_plain.log profiles/r2_ncu_full_summary.csv

The second argument is the plain run's log ("profiled launch N: label" lines give the case names in launch order)."""
import csv
import re
import sys

raw, log, out = sys.argv[1:4]
labels = {}
for line in open(log):
    m = re.match(r"profiled launch\s+(\d+): (.*)", line)
    if m:
        labels[int(m.group(1))] = m.group(2).strip()
    m = re.match(r"profiled launches\s+(\d+),\s+(\d+): (.*)", line)
    if m:
        labels[int(m.group(1))] = m.group(3).strip() + " (fwd)"
        labels[int(m.group(2))] = m.group(3).strip() + " (bwd)"
KEEP = ["launch__grid_size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum"]
rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
head, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(head)}
cols = [k for k in KEEP if k in idx]
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["case", "ID", "Kernel Name"] + ["%s[%s]" % (c, units[idx[c]]) if units[idx[c]] else c for c in cols])
    for n, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("acg::", "").replace("(anonymous namespace)::", "")
        w.writerow([labels.get(n, ""), r[idx["ID"]], name] + [r[idx[c]] for c in cols])
print("wrote", out, len(data), "launches")
