#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --method accurate"
timeout 300 $CMD > gpurun_out/plain_full.log 2> gpurun_out/plain_full.err &&
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|dec_fused|conv1_bn|preprocess_pack|ctc_greedy|encoder_attention|pool_pos|ln_chain" -s 395 -c 40 -o /tmp/prof_r1_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "== ncu rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/prof_r1_full.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_raw.csv 2>/dev/null; ls -la gpurun_out/r01_ncu_full_raw.csv /tmp/prof_r1_full.ncu-rep
