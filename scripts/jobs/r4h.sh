#!/bin/bash
# round 2, session 3, GPU call H: first layers on the product library, staged vs direct stores
mkdir -p gpurun_out
timeout 300 python scripts/first_layer_ab.py ACG_EPI_DIRECT g/conv1,d/conv1,g/tconv3 256 > gpurun_out/r4h_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/r4h_ab.log
tail -n 6 gpurun_out/r4h_ab.log
exit 0
