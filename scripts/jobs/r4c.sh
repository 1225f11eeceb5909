#!/bin/bash
# round 2, session 3, GPU call C: small-K kernel with warp-local transposed stores (parity, A/B); DNA copy issue A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_determinism_gpu.py -m gpu -q -x -k "small_k or determinis or reproduc" > gpurun_out/r4c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4c_tests.log
tail -n 4 gpurun_out/r4c_tests.log
timeout 300 python scripts/first_layer_ab.py ACG_EPI_DIRECT g/conv1,d/conv1 256 > gpurun_out/r4c_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/r4c_ab.log
cat gpurun_out/r4c_ab.log | tail -n 8
for lanes in 32 1 32 1; do
  echo "issue lanes $lanes"
  ACG_DNA_ISSUE_LANES=$lanes timeout 200 python scripts/dna_sweep.py quick 2>&1 | tail -n 4
done > gpurun_out/r4c_dna_ab.log 2>&1
cat gpurun_out/r4c_dna_ab.log
exit 0
