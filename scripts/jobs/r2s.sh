#!/bin/bash
# round 2, GPU call S2: weight-gradient whole-wave splits, 1 or 2 waves
mkdir -p gpurun_out
for w in 1 2; do
  ACG_WGRAD_WAVES=$w timeout 200 python scripts/probe_r2.py 256 > gpurun_out/r2s_probe_w$w.log 2>&1
  echo "== waves $w"; sed -n 3,20p gpurun_out/r2s_probe_w$w.log | awk '{print $1, $(NF-1), $NF}'
  ACG_WGRAD_WAVES=$w timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
done
exit 0
