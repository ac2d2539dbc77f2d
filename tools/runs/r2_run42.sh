#!/bin/bash
# round 2, call 42: CTC logits kernel, final form: tests, HBM exhibit, DRAM bytes (ncu) at 8192 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_beam_gpu.py tests/test_api_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python tools/bench_hbm_kernels.py > gpurun_out/r2_42_hbm.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_42_hbm.json'))
print({k:(round(v['ms'],4), round(v['frac_of_hbm_peak'],3)) for k,v in d.items()})"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"ctc_greedy" -s 1 -c 1 --csv --log-file gpurun_out/r2_42_ctc_ncu.csv python tools/hbm_once.py 8192 > gpurun_out/r2_42_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2_42_ctc_ncu.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -7
