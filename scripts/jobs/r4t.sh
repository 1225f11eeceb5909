#!/bin/bash
# round 2, session 3, GPU call T: what the fused backward reduction costs per data-gradient launch
mkdir -p gpurun_out
timeout 200 python scripts/fused_red_cost.py 256 > gpurun_out/r4t_fused_red.log 2>&1
echo "rc=$?" >> gpurun_out/r4t_fused_red.log
tail -n 10 gpurun_out/r4t_fused_red.log
exit 0
