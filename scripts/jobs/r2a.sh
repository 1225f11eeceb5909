#!/bin/bash
# round 2, GPU call A: regression tests, new un-gated parity tests, probes, bench, ncu of the current kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_fullstep_parity_gpu.py > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 900 python -m pytest tests/test_fullstep_parity_gpu.py -m gpu -q > gpurun_out/r2a_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2a_parity.log
timeout 600 python scripts/probe_r2.py 256 > gpurun_out/r2a_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2a_probe.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 300 python scripts/ncu_kernels.py > gpurun_out/r2a_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2a_kernels \
    python scripts/ncu_kernels.py > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2a_ncu.log
ls -la gpurun_out | tail -20
tail -5 gpurun_out/r2a_pytest.log gpurun_out/r2a_parity.log
