#!/bin/bash
# round 2, GPU call D: lean warp-uniform MMA issue loop, torch.library ops
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_tc_gpu.py tests/test_torch_ops_gpu.py tests/test_feeder_gpu.py -m gpu -q > gpurun_out/r2d_new.log 2>&1
echo "new rc=$?" >> gpurun_out/r2d_new.log
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2d_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2d_probe.log
timeout 600 python scripts/step_time.py 256 30 > gpurun_out/r2d_step.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_conv_halo_gpu.py --deselect tests/test_conv_tc_gpu.py --deselect tests/test_torch_ops_gpu.py --deselect tests/test_feeder_gpu.py > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -n 6 gpurun_out/r2d_new.log gpurun_out/r2d_pytest.log gpurun_out/r2d_step.log
exit 0
