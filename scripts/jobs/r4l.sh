#!/bin/bash
# round 2, session 3, GPU call L (2 GPUs): data-parallel parity tests + one bench run after this session's kernel changes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -x -k "not 4-peer" > gpurun_out/r4l_peer_2gpu.log 2>&1
echo "peer rc=$?" >> gpurun_out/r4l_peer_2gpu.log
tail -n 4 gpurun_out/r4l_peer_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r4l_bench_2gpu.json 2> gpurun_out/r4l_bench_2gpu.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r4l_bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","step_ms_median","step_ms_max")}, "e2e", d["e2e"]["value"], "strong", d.get("strong_scaling"))
PY
exit 0
