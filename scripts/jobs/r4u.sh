#!/bin/bash
# round 2, session 3, GPU call U: fused backward reduction with batched z loads -- per launch, parity, step time
mkdir -p gpurun_out
timeout 200 python scripts/fused_red_cost.py 256 > gpurun_out/r4u_fused_red.log 2>&1
tail -n 7 gpurun_out/r4u_fused_red.log
timeout 200 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -k "fused_bwd" > gpurun_out/r4u_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r4u_tests.log
tail -n 2 gpurun_out/r4u_tests.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_FUSE_BWD_REDUCE=halo timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
exit 0
