#!/bin/bash
# round 2, GPU call V: traversal direction of the streaming batch-norm passes (L2 reuse)
mkdir -p gpurun_out
for r in 0 1 2 4 3 5; do
  ACG_EW_REV=$r timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
done
ACG_EW_REV=3 timeout 300 python -m pytest tests/test_elementwise_gpu.py tests/test_networks_gpu.py -m gpu -q -x 2>&1 | tail -n 3
exit 0
