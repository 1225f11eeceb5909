#!/bin/bash
# round 2, GPU call R2: weight-gradient kernel with the x operand by TMA (bounded: -x and short timeouts)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_halo_gpu.py -m gpu -q -x -k "wgrad or first_layers or full_size" > gpurun_out/r2r_wgrad.log 2>&1
rc=$?
echo "wgrad rc=$rc" >> gpurun_out/r2r_wgrad.log
tail -n 5 gpurun_out/r2r_wgrad.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 python -m pytest tests/test_fullstep_parity_gpu.py tests/test_networks_gpu.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2r_tests.log
timeout 200 python scripts/step_time.py 256 30 > gpurun_out/r2r_step.log 2>&1
timeout 300 python scripts/probe_r2.py 256 > gpurun_out/r2r_probe.log 2>&1
tail -n 5 gpurun_out/r2r_tests.log; tail -n 2 gpurun_out/r2r_step.log; head -22 gpurun_out/r2r_probe.log
exit 0
