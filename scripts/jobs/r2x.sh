#!/bin/bash
# round 2, GPU call X: final single-GPU evidence -- full suite, benches (all configs), ncu launch list + --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2x_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -n 4 gpurun_out/r2x_tests.log
timeout 600 python bench.py > gpurun_out/r2x_bench_1gpu.json 2> gpurun_out/r2x_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --config wass_rmsprop --steps 30 > gpurun_out/r2x_bench_wass.json 2> gpurun_out/r2x_bench_wass.err; echo "wass rc=$?"
timeout 300 python bench.py --config direct_rollout --steps 30 > gpurun_out/r2x_bench_rollout.json 2> gpurun_out/r2x_bench_rollout.err; echo "rollout rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; echo "reference rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1; tail -n 1 gpurun_out/r2x_smoke.log
timeout 300 python scripts/probe_r2.py 256 > gpurun_out/r2x_probe.log 2>&1
timeout 600 python scripts/ablate.py 256 > gpurun_out/r2x_ablate.log 2>&1
timeout 300 python scripts/one_iter.py > gpurun_out/r2x_one_iter_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2_iter_launches.csv python scripts/one_iter.py > gpurun_out/r2x_one_iter_ncu.log 2>&1
echo "one_iter ncu rc=$?"
timeout 300 python scripts/ncu_kernels.py > gpurun_out/r2x_kernels_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_kernels \
    python scripts/ncu_kernels.py > gpurun_out/r2x_kernels_ncu.log 2>&1
echo "kernels ncu rc=$?"
for f in gpurun_out/r2x_bench_1gpu.json gpurun_out/r2x_bench_wass.json gpurun_out/r2x_bench_rollout.json gpurun_out/r2x_bench_reference.json; do cut -c1-330 $f; echo; done
ls -la gpurun_out | grep "r2_"
exit 0
