#!/bin/bash
# round 2, session 3, GPU call G: full GPU test suite, default bench line, per-layer table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r4g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4g_tests.log
tail -n 14 gpurun_out/r4g_tests.log
timeout 300 python bench.py > gpurun_out/r4g_bench.json 2> gpurun_out/r4g_bench.err
echo "bench rc=$?"
cut -c1-700 gpurun_out/r4g_bench.json
timeout 300 python scripts/probe_r2.py 256 > gpurun_out/r4g_probe.log 2>&1
sed -n 1,22p gpurun_out/r4g_probe.log
exit 0
