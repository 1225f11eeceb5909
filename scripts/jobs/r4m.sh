#!/bin/bash
# round 2, session 3, GPU call M: weight-ring depth sensitivity of the halo kernel
mkdir -p gpurun_out
timeout 300 python scripts/halo_ring_sweep.py 256 > gpurun_out/r4m_ring.log 2>&1
echo "rc=$?" >> gpurun_out/r4m_ring.log
tail -n 16 gpurun_out/r4m_ring.log
exit 0
