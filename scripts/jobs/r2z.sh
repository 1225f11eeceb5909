#!/bin/bash
# round 2, GPU call Z (8 GPUs): BASELINE configs[3] (wass + rmsprop) and configs[4] (direct generator rollout, B=512) data parallel
mkdir -p gpurun_out
for c in wass_rmsprop direct_rollout; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --config $c --steps 20 --warmup 3 > gpurun_out/r2z_bench_8gpu_$c.json 2> gpurun_out/r2z_bench_8gpu_$c.err
  echo "$c rc=$?"; tail -n 1 gpurun_out/r2z_bench_8gpu_$c.json | cut -c1-420; tail -n 2 gpurun_out/r2z_bench_8gpu_$c.err | cut -c1-300
done
exit 0
