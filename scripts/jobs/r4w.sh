#!/bin/bash
# round 2, session 3, GPU call W (2 GPUs): data-parallel parity tests on the final build (staged epilogue + ticketed finalize + peer exchange)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_peer_gpu.py -m gpu -q -rs -x -k "not 4-peer" > gpurun_out/r4w_peer_2gpu.log 2>&1
echo "peer rc=$?" >> gpurun_out/r4w_peer_2gpu.log
tail -n 4 gpurun_out/r4w_peer_2gpu.log
exit 0
