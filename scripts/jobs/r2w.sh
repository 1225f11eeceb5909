#!/bin/bash
# round 2, GPU call W: weight-gradient TMA path for padded channel atoms + balanced DNA bands
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "wgrad or first_layers" > gpurun_out/r2w_wgrad.log 2>&1
rc=$?
echo "wgrad rc=$rc" >> gpurun_out/r2w_wgrad.log
tail -n 4 gpurun_out/r2w_wgrad.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python -m pytest tests/test_fullstep_parity_gpu.py tests/test_dna_gpu.py -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -n 3 gpurun_out/r2w_tests.log
timeout 200 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 300 python scripts/probe_r2.py 256 > gpurun_out/r2w_probe.log 2>&1
sed -n 3,20p gpurun_out/r2w_probe.log | awk '{print $1, $(NF-1), $NF}'
timeout 200 python scripts/dna_bench.py > gpurun_out/r2w_dna.json 2>&1; cat gpurun_out/r2w_dna.json | tr -d '\n' | cut -c1-700
exit 0
