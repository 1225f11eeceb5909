#!/bin/bash
# round 2, GPU call U (2 GPUs): fused SyncBN exchange, tagged-cell protocol
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -x -k "not 4-peer" > gpurun_out/r2u_peer.log 2>&1
rc=$?
echo "peer rc=$rc" >> gpurun_out/r2u_peer.log
tail -n 8 gpurun_out/r2u_peer.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2u_bench_2gpu.json; tail -n 2 gpurun_out/r2u_bench_2gpu.err
ACG_DP_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2u_bench_2gpu_unfused.json 2> gpurun_out/r2u_bench_2gpu_unfused.err
echo "bench (unfused) rc=$?"; cut -c1-300 gpurun_out/r2u_bench_2gpu_unfused.json
exit 0
