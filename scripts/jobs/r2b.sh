#!/bin/bash
# round 2, GPU call B: new halo kernel (both forms), elect-based MMA issue everywhere, trainer restructure
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_tc_gpu.py -m gpu -q -x > gpurun_out/r2b_halo.log 2>&1
echo "halo rc=$?" >> gpurun_out/r2b_halo.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_conv_halo_gpu.py --deselect tests/test_conv_tc_gpu.py > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2b_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2b_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" >> gpurun_out/r2b_bench.err
ACG_OVERLAP_G_FWD=0 timeout 600 python scripts/step_time.py 256 20 > gpurun_out/r2b_step_nooverlap.log 2>&1
timeout 600 python scripts/step_time.py 256 20 > gpurun_out/r2b_step.log 2>&1
tail -n 5 gpurun_out/r2b_halo.log gpurun_out/r2b_pytest.log gpurun_out/r2b_step.log gpurun_out/r2b_step_nooverlap.log
exit 0
