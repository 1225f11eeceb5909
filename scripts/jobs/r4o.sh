#!/bin/bash
# round 2, session 3, GPU call O: pixel-major kernel -- staged epilogue with column-wise moments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_px_gpu.py tests/test_conv_halo_gpu.py tests/test_determinism_gpu.py tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py tests/test_conv_tc_gpu.py -m gpu -q -x > gpurun_out/r4o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4o_tests.log
tail -n 4 gpurun_out/r4o_tests.log
timeout 200 python scripts/stats_cost.py 256 > gpurun_out/r4o_stats_cost.log 2>&1
tail -n 17 gpurun_out/r4o_stats_cost.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_EPI_DIRECT=1 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
