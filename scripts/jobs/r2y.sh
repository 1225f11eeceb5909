#!/bin/bash
# round 2, GPU call Y (N GPUs): data-parallel tests and bench at N = $1
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then K="not 4-peer"; else K="4-peer"; fi
if [ "$N" != "8" ]; then
  timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -x -k "$K" > gpurun_out/r2y_peer_${N}gpu.log 2>&1
  echo "peer rc=$?" >> gpurun_out/r2y_peer_${N}gpu.log
  tail -n 4 gpurun_out/r2y_peer_${N}gpu.log
fi
for i in 1 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2y_bench_${N}gpu_$i.json 2> gpurun_out/r2y_bench_${N}gpu_$i.err
  echo "bench $i rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_${N}gpu_$i.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","step_ms_median","step_ms_max","step_ms_argmax_rank0")}, "e2e", d["e2e"]["value"], "strong", d.get("strong_scaling"))
PY
done
exit 0
