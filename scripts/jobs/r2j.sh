#!/bin/bash
# round 2, GPU call J (4 GPUs): world-4 data-parallel parity + three N=4 bench repeats (per-step max exposes stalls)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -v -k "4-peer or world4 or 4" > gpurun_out/r2j_peer4.log 2>&1
echo "peer4 rc=$?" >> gpurun_out/r2j_peer4.log
for i in 1 2 3; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r2j_bench_4gpu_$i.json 2> gpurun_out/r2j_bench_4gpu_$i.err
  echo "bench $i rc=$?" >> gpurun_out/r2j_bench_4gpu_$i.err
done
tail -n 8 gpurun_out/r2j_peer4.log
exit 0
