#!/bin/bash
# round 2, GPU call AA: action concat in the activation pass + bench line with the conv-only graph roofline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_networks_gpu.py tests/test_fullstep_parity_gpu.py -m gpu -q -x > gpurun_out/r3a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r3a_tests.log
tail -n 3 gpurun_out/r3a_tests.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/r3a_bench_1gpu.json 2> gpurun_out/r3a_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"])
PY
tail -n 3 gpurun_out/r3a_bench_1gpu.err
exit 0
