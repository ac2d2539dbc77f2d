#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/e2e_diag.log 2>&1
nproc >> gpurun_out/e2e_diag.log
timeout 300 python tools/e2e_diag.py >> gpurun_out/e2e_diag.log 2>&1; echo "rc=$?"
cat gpurun_out/e2e_diag.log
