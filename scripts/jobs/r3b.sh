#!/bin/bash
# round 2, GPU call AB: first layers in the halo kernel's pixel-pair mode
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_conv_halo_gpu.py -m gpu -q -x -k "pixel_pair" > gpurun_out/r3b_pair.log 2>&1
rc=$?
echo "pair rc=$rc" >> gpurun_out/r3b_pair.log
tail -n 12 gpurun_out/r3b_pair.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 python -m pytest tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py tests/test_determinism_gpu.py tests/test_trainer_gpu.py -m gpu -q -x > gpurun_out/r3b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r3b_tests.log
tail -n 3 gpurun_out/r3b_tests.log
timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
ACG_PAIR_FIRST=0 timeout 120 python scripts/step_time.py 256 30 2>&1 | tail -n 1
timeout 300 python scripts/probe_r2.py 256 > gpurun_out/r3b_probe.log 2>&1
sed -n 3,20p gpurun_out/r3b_probe.log | awk '{print $1, $5, $6}'
exit 0
