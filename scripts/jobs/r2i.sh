#!/bin/bash
# round 2, GPU call I: ncu evidence of the current kernels (launch list with DRAM bytes of one iteration, --set full of each
# conv kernel variant)
mkdir -p gpurun_out
timeout 300 python scripts/one_iter.py > gpurun_out/r2i_one_iter_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2_iter_launches.csv python scripts/one_iter.py > gpurun_out/r2i_one_iter_ncu.log 2>&1
echo "one_iter ncu rc=$?" >> gpurun_out/r2i_one_iter_ncu.log
timeout 300 python scripts/ncu_kernels.py > gpurun_out/r2i_kernels_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_kernels \
    python scripts/ncu_kernels.py > gpurun_out/r2i_kernels_ncu.log 2>&1
echo "kernels ncu rc=$?" >> gpurun_out/r2i_kernels_ncu.log
ls -la gpurun_out | grep r2_; tail -n 3 gpurun_out/r2i_one_iter_ncu.log gpurun_out/r2i_kernels_ncu.log
exit 0
