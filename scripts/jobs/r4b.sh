#!/bin/bash
# round 2, session 3, GPU call B: DNA -- copies issued by several lanes; ring depth / grid sweep (probe library)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dna_gpu.py -m gpu -q -x > gpurun_out/r4b_dna_tests.log 2>&1
echo "dna tests rc=$?" >> gpurun_out/r4b_dna_tests.log
tail -n 4 gpurun_out/r4b_dna_tests.log
timeout 400 python scripts/dna_sweep.py > gpurun_out/r4b_dna_sweep.log 2>&1
echo "sweep rc=$?" >> gpurun_out/r4b_dna_sweep.log
cat gpurun_out/r4b_dna_sweep.log
exit 0
