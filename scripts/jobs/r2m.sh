#!/bin/bash
# round 2, GPU call M: 8-producer-warp weight-gradient kernel + reproducible fused moments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py tests/test_conv_halo_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2m_tests.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2m_step.log 2>&1
ACG_DETERMINISTIC=0 timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2m_step_atomics.log 2>&1
timeout 300 python scripts/determinism_probe.py > gpurun_out/r2m_determinism.log 2>&1
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2m_probe.log 2>&1
tail -n 6 gpurun_out/r2m_tests.log; tail -n 3 gpurun_out/r2m_step.log gpurun_out/r2m_step_atomics.log; cat gpurun_out/r2m_determinism.log; head -30 gpurun_out/r2m_probe.log
exit 0
