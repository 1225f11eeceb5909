#!/bin/bash
# round 2, GPU call Q: px split prefetch + critical-path ablation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_px_gpu.py tests/test_determinism_gpu.py -m gpu -q > gpurun_out/r2q_px.log 2>&1
echo "px rc=$?" >> gpurun_out/r2q_px.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2q_step.log 2>&1
timeout 900 python scripts/ablate.py 256 > gpurun_out/r2q_ablate.log 2>&1
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2q_probe.log 2>&1
tail -n 5 gpurun_out/r2q_px.log; tail -n 2 gpurun_out/r2q_step.log; cat gpurun_out/r2q_ablate.log; head -22 gpurun_out/r2q_probe.log
exit 0
