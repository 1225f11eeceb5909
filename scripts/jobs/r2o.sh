#!/bin/bash
# round 2, GPU call O: fixed-point limb accumulators for the fused moments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py tests/test_conv_halo_gpu.py tests/test_golden_gpu.py -m gpu -q > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2o_tests.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2o_step.log 2>&1
ACG_DETERMINISTIC=0 timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2o_step_atomics.log 2>&1
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2o_step2.log 2>&1
tail -n 8 gpurun_out/r2o_tests.log; tail -n 2 gpurun_out/r2o_step.log gpurun_out/r2o_step_atomics.log gpurun_out/r2o_step2.log
exit 0
