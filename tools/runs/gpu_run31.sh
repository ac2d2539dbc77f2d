#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stem12" > gpurun_out/pytest_stem12.log 2>&1
echo "== stem12 rc=$?"; grep -E "passed|failed|FAILED|Error|max err|assert" gpurun_out/pytest_stem12.log | tail -8
