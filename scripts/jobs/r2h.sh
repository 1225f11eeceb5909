#!/bin/bash
# round 2, GPU call H: 8x8 interleaved halo mode, selective fused reduction
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_halo_gpu.py -m gpu -q > gpurun_out/r2h_conv.log 2>&1
echo "conv rc=$?" >> gpurun_out/r2h_conv.log
ACG_FUSE_BWD_REDUCE=0 timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2h_step_nofuse.log 2>&1
timeout 300 python scripts/step_time.py 256 30 > gpurun_out/r2h_step.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_conv_halo_gpu.py --deselect tests/test_conv_tc_gpu.py > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2h_probe.log 2>&1
tail -n 4 gpurun_out/r2h_conv.log gpurun_out/r2h_step_nofuse.log gpurun_out/r2h_step.log gpurun_out/r2h_pytest.log
exit 0
