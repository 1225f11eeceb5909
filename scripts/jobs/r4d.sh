#!/bin/bash
# round 2, session 3, GPU call D: small-K kernel with eight producer warps
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py -m gpu -q -x -k "small_k or determinis or reproduc" > gpurun_out/r4d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4d_tests.log
tail -n 4 gpurun_out/r4d_tests.log
timeout 300 python scripts/first_layer_ab.py ACG_EPI_DIRECT g/conv1,d/conv1 256 > gpurun_out/r4d_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/r4d_ab.log
cat gpurun_out/r4d_ab.log | tail -n 8
exit 0
