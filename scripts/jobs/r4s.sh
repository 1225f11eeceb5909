#!/bin/bash
# round 2, session 3, GPU call S: ncu launch list (time + DRAM bytes) of ONE iteration in the final state
mkdir -p gpurun_out
timeout 300 python scripts/one_iter.py > gpurun_out/r4s_one_iter_plain.log 2>&1
echo "plain rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2s3_iter_launches.csv python scripts/one_iter.py > gpurun_out/r4s_one_iter_ncu.log 2>&1
echo "one_iter ncu rc=$?"
wc -l gpurun_out/r2s3_iter_launches.csv
exit 0
