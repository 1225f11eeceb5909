#!/bin/bash
# round 2, session 3, GPU call I: stage knock-out of the pixel-major kernel
mkdir -p gpurun_out
timeout 300 python scripts/px_knockout.py 256 > gpurun_out/r4i_px_knockout.log 2>&1
echo "rc=$?" >> gpurun_out/r4i_px_knockout.log
tail -n 12 gpurun_out/r4i_px_knockout.log
exit 0
