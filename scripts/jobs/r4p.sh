#!/bin/bash
# round 2, session 3, GPU call P: halo kernel -- staged stores also for the data gradients (no moments)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_tc_gpu.py -m gpu -q -x > gpurun_out/r4p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4p_tests.log
tail -n 3 gpurun_out/r4p_tests.log
timeout 300 python scripts/first_layer_ab.py ACG_EPI_DIRECT g/conv2,g/tconv2,g/tconv3,g/tconv4,d/conv2,d/conv3 256 > gpurun_out/r4p_ab.log 2>&1
tail -n 8 gpurun_out/r4p_ab.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
