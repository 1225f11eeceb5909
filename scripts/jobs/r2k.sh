#!/bin/bash
# round 2, GPU call K: N=256-column weight-gradient kernel (tests, layer table, step time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -k "wgrad or first_layers" > gpurun_out/r2k_wgrad.log 2>&1
echo "wgrad rc=$?" >> gpurun_out/r2k_wgrad.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2k_step.log 2>&1
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2k_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -n 4 gpurun_out/r2k_wgrad.log gpurun_out/r2k_step.log gpurun_out/r2k_pytest.log; head -22 gpurun_out/r2k_probe.log
exit 0
