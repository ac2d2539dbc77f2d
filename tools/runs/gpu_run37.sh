#!/bin/bash
# ncu evidence for the current build: launch list of one bench step + full-set capture of the top kernels
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_v4.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_fast_v4.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encoder_block_kernel -s 4 -c 2 -f -o gpurun_out/ncu_eb \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_eb.log 2>&1; echo "eb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 80 -c 14 -f -o gpurun_out/ncu_gemm \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"encoder_attention_kernel|conv1_bn_silu_kernel|preprocess_pack_kernel|ctc_greedy" -s 20 -c 10 -f -o gpurun_out/ncu_misc \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_misc.log 2>&1; echo "misc rc=$?"
ls -la gpurun_out/*.ncu-rep
