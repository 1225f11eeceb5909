#!/bin/bash
# round 2, GPU call E (2 GPUs): data-parallel parity tests + 2-GPU bench (weak + strong scaling)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -v > gpurun_out/r2e_peer.log 2>&1
echo "peer rc=$?" >> gpurun_out/r2e_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
echo "bench rc=$?" >> gpurun_out/r2e_bench_2gpu.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err
tail -n 15 gpurun_out/r2e_peer.log; tail -n 3 gpurun_out/r2e_bench_2gpu.err
exit 0
