#!/bin/bash
# round 2, GPU call L: weight-gradient kernel with eight producer warps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -k "wgrad or first_layers" > gpurun_out/r2l_wgrad.log 2>&1
echo "wgrad rc=$?" >> gpurun_out/r2l_wgrad.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2l_step.log 2>&1
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2l_probe.log 2>&1
tail -n 4 gpurun_out/r2l_wgrad.log gpurun_out/r2l_step.log; head -22 gpurun_out/r2l_probe.log
exit 0
