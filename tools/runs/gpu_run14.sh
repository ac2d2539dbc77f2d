#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_timing.py > gpurun_out/gemm_timing.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_timing.log
