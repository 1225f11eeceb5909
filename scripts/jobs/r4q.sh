#!/bin/bash
# round 2, session 3, GPU call Q: batch-norm backward reduction in the staged data-gradient epilogue (halo kernel)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_halo_gpu.py -m gpu -q -x -k "fused_bwd or halo" > gpurun_out/r4q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4q_tests.log
tail -n 3 gpurun_out/r4q_tests.log
ACG_FUSE_BWD_REDUCE=halo timeout 400 python -m pytest tests/test_fullstep_parity_gpu.py tests/test_trainer_gpu.py tests/test_networks_gpu.py -m gpu -q -x > gpurun_out/r4q_tests_fused.log 2>&1
echo "fused tests rc=$?" >> gpurun_out/r4q_tests_fused.log
tail -n 3 gpurun_out/r4q_tests_fused.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_FUSE_BWD_REDUCE=halo timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_FUSE_BWD_REDUCE=halo timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
