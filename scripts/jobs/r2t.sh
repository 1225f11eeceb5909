#!/bin/bash
# round 2, GPU call T: full GPU suite + default bench (mid-round evidence)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2t_tests.log
tail -n 6 gpurun_out/r2t_tests.log
timeout 600 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
echo "bench rc=$?"; cat gpurun_out/r2t_bench.json; tail -n 3 gpurun_out/r2t_bench.err
exit 0
