#!/bin/bash
# round 2, GPU call G: fused batch-norm backward reduction in the halo epilogue (tests + A/B of the step time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_halo_gpu.py -m gpu -q > gpurun_out/r2g_conv.log 2>&1
echo "conv rc=$?" >> gpurun_out/r2g_conv.log
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2g_step_sep.log 2>&1
ACG_FUSE_BWD_REDUCE=1 timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2g_step_fused.log 2>&1
ACG_FUSE_BWD_REDUCE=1 timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py -m gpu -q > gpurun_out/r2g_fused_tests.log 2>&1
echo "fused tests rc=$?" >> gpurun_out/r2g_fused_tests.log
tail -n 4 gpurun_out/r2g_conv.log gpurun_out/r2g_step_sep.log gpurun_out/r2g_step_fused.log gpurun_out/r2g_fused_tests.log
exit 0
