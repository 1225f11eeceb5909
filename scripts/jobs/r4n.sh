#!/bin/bash
# round 2, session 3, GPU call N: halo kernel -- staged epilogue with column-wise moments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_halo_gpu.py tests/test_determinism_gpu.py tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py -m gpu -q -x > gpurun_out/r4n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4n_tests.log
tail -n 6 gpurun_out/r4n_tests.log
timeout 200 python scripts/stats_cost.py 256 > gpurun_out/r4n_stats_cost.log 2>&1
tail -n 17 gpurun_out/r4n_stats_cost.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_EPI_DIRECT=1 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_EPI_DIRECT=1 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
