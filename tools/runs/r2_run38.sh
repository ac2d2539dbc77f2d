#!/bin/bash
# round 2, call 38: DRAM bytes (ncu) of the two byte-bound kernels at 8192 lines, next to their algorithmic bytes
mkdir -p gpurun_out
timeout 300 python tools/hbm_once.py 8192 > gpurun_out/r2_38_alg.txt 2>&1 || { tail -5 gpurun_out/r2_38_alg.txt; exit 1; }
cat gpurun_out/r2_38_alg.txt | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"preprocess_pack|crop_sum|ctc_greedy" -s 3 -c 3 --csv --log-file gpurun_out/r2_38_hbm_ncu.csv python tools/hbm_once.py 8192 > gpurun_out/r2_38_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2_38_hbm_ncu.csv | cut -d, -f5,13,15 | head -30
