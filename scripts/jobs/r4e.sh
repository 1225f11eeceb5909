#!/bin/bash
# round 2, session 3, GPU call E: stage knock-out of the small-K kernel
mkdir -p gpurun_out
timeout 300 python scripts/smallk_knockout.py 256 > gpurun_out/r4e_knockout.log 2>&1
echo "rc=$?" >> gpurun_out/r4e_knockout.log
cat gpurun_out/r4e_knockout.log | tail -n 8
exit 0
