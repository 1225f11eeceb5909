#!/bin/bash
# round 2, session 3, GPU call K: batch-norm finalize inside the activation pass (no ticket in the conv launches)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_determinism_gpu.py tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py tests/test_trainer_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r4k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4k_tests.log
tail -n 12 gpurun_out/r4k_tests.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_RAW_MOMENTS=0 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_RAW_MOMENTS=0 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
