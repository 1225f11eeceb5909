#!/bin/bash
# round 2, session 3, GPU call J: cost of the fused batch-norm moments per launch
mkdir -p gpurun_out
timeout 300 python scripts/stats_cost.py 256 > gpurun_out/r4j_stats_cost.log 2>&1
echo "rc=$?" >> gpurun_out/r4j_stats_cost.log
tail -n 22 gpurun_out/r4j_stats_cost.log
exit 0
