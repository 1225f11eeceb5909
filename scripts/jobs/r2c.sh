#!/bin/bash
# round 2, GPU call C: deeper weight ring in the halo kernel, fixed tests, new bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_tc_gpu.py -m gpu -q > gpurun_out/r2c_halo.log 2>&1
echo "halo rc=$?" >> gpurun_out/r2c_halo.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_conv_halo_gpu.py --deselect tests/test_conv_tc_gpu.py > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2c_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2c_probe.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?" >> gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --config wass_rmsprop > gpurun_out/r2c_bench_wass.json 2> gpurun_out/r2c_bench_wass.err
timeout 600 python bench.py --steps 10 --warmup 3 --config direct_rollout > gpurun_out/r2c_bench_rollout.json 2> gpurun_out/r2c_bench_rollout.err
timeout 600 python scripts/ablate.py 256 > gpurun_out/r2c_ablate.log 2>&1
tail -n 4 gpurun_out/r2c_halo.log gpurun_out/r2c_pytest.log gpurun_out/r2c_bench.err gpurun_out/r2c_bench_wass.err gpurun_out/r2c_bench_rollout.err
exit 0
