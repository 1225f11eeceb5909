#!/bin/bash
# round 2, session 3, GPU call R: final single-GPU evidence -- full suite, smoke, bench lines, layer table, ablation
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4r_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4r_tests.log
tail -n 3 gpurun_out/r4r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4r_smoke.log 2>&1; tail -n 1 gpurun_out/r4r_smoke.log
timeout 400 python bench.py > gpurun_out/r4r_bench_1gpu.json 2> gpurun_out/r4r_bench_1gpu.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/r4r_bench_1gpu.json
timeout 300 python bench.py --config wass_rmsprop --steps 30 > gpurun_out/r4r_bench_wass.json 2> gpurun_out/r4r_bench_wass.err; echo "wass rc=$?"
timeout 300 python bench.py --config direct_rollout --steps 30 > gpurun_out/r4r_bench_rollout.json 2> gpurun_out/r4r_bench_rollout.err; echo "rollout rc=$?"
timeout 200 python scripts/probe_r2.py 256 > gpurun_out/r4r_probe.log 2>&1
sed -n 3,20p gpurun_out/r4r_probe.log | awk '{print $1, $5, $8, $11}' | tr '\n' ';'
timeout 400 python scripts/ablate.py 256 > gpurun_out/r4r_ablate.log 2>&1
sed -n 1,16p gpurun_out/r4r_ablate.log
exit 0
