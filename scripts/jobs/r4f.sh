#!/bin/bash
# round 2, session 3, GPU call F: small-K kernel -- deeper ring, trimmed last K block, N-dependent producer count
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py -m gpu -q -x -k "small_k or determinis or reproduc" > gpurun_out/r4f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4f_tests.log
tail -n 4 gpurun_out/r4f_tests.log
timeout 300 python scripts/smallk_knockout.py 256 > gpurun_out/r4f_knockout.log 2>&1
echo "rc=$?" >> gpurun_out/r4f_knockout.log
cat gpurun_out/r4f_knockout.log | tail -n 8
exit 0
