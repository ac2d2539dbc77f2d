#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --width-mode parity"
timeout 300 $CMD > gpurun_out/plain_attn.log 2> gpurun_out/plain_attn.err &&
timeout 600 ncu --set full --clock-control none -k regex:"encoder_attention" -s 8 -c 2 -o /tmp/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "== ncu rc=$?"
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/r01_ncu_attn_raw.csv 2>/dev/null
ncu -i /tmp/prof_attn.ncu-rep --page details --csv > gpurun_out/r01_ncu_attn_details.csv 2>/dev/null
ls -la gpurun_out/r01_ncu_attn_*.csv
