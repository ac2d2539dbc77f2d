#!/bin/bash
# round 2, call 21: role-level cycle breakdown of the conv / QKV launches
mkdir -p gpurun_out
timeout 300 python tools/gemm_timing.py > gpurun_out/r2_21_gemm_timing.txt 2>&1; echo rc=$?; cat gpurun_out/r2_21_gemm_timing.txt
