#!/bin/bash
# round 2, GPU call P: pixel-major kernel for small feature maps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_px_gpu.py -m gpu -q > gpurun_out/r2p_px.log 2>&1
echo "px rc=$?" >> gpurun_out/r2p_px.log
timeout 1200 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py tests/test_networks_gpu.py tests/test_fullstep_parity_gpu.py -m gpu -q > gpurun_out/r2p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2p_tests.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2p_step.log 2>&1
ACG_NO_PX=1 timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2p_step_nopx.log 2>&1
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2p_probe.log 2>&1
grep -v "^$" gpurun_out/r2p_px.log | tail -n 40; tail -n 8 gpurun_out/r2p_tests.log; tail -n 2 gpurun_out/r2p_step.log gpurun_out/r2p_step_nopx.log; head -22 gpurun_out/r2p_probe.log
exit 0
