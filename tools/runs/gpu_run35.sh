#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "encoder_block" > gpurun_out/pytest_eb.log 2>&1
echo "== encoder_block rc=$?"; grep -E "passed|failed|FAILED|Error|assert|timeout|kiri:" gpurun_out/pytest_eb.log | tail -12
