#!/bin/bash
# round 2, GPU call F: compute-sanitizer memcheck over every kernel family at small shapes (+ DNA prologue change)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dna_gpu.py tests/test_golden_gpu.py -m gpu -q > gpurun_out/r2f_dna.log 2>&1
echo "dna rc=$?" >> gpurun_out/r2f_dna.log
timeout 300 python scripts/sanitize_small.py > gpurun_out/r2f_plain.log 2>&1
echo "plain rc=$?" >> gpurun_out/r2f_plain.log
if grep -q "sanitize_small: done" gpurun_out/r2f_plain.log; then
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2f_memcheck.log 2>&1
  echo "memcheck rc=$?" >> gpurun_out/r2f_memcheck.log
fi
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -n 5 gpurun_out/r2f_dna.log gpurun_out/r2f_plain.log; tail -n 12 gpurun_out/r2f_memcheck.log
exit 0
