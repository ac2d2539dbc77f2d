#!/bin/bash
# First on-device bring-up: every kernel group in its own process so one fault cannot mask the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/diag_gemm.py > gpurun_out/diag_gemm.log 2>&1; echo "diag rc=$?" >> gpurun_out/diag_gemm.log
for k in ctc preprocess gemm conv3x3 conv1 pool_pos attention; do
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA -k "$k" > gpurun_out/pytest_$k.log 2>&1
  echo "== $k rc=$?"; tail -3 gpurun_out/pytest_$k.log
done
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA > gpurun_out/pytest_engine.log 2>&1
echo "== engine rc=$?"; tail -5 gpurun_out/pytest_engine.log
cat gpurun_out/diag_gemm.log | tail -40
