#!/bin/bash
# round 2, call 65: ncu full-set capture of the three kernels changed last (pool, QKV GEMM, attention) in the final build
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|encoder_attention_kernel|pool_pos_ln_kernel|ln_chain_kernel" \
    -s 31 -c 3 -f -o gpurun_out/r02_final_ncu_qkv_attn python bench.py --steps 2 --warmup 3 > gpurun_out/r02_final_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_final_ncu_qkv_attn.ncu-rep
