#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA -k "preprocess" > gpurun_out/pytest_preprocess.log 2>&1
echo "== preprocess rc=$?"; tail -4 gpurun_out/pytest_preprocess.log
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA > gpurun_out/pytest_engine.log 2>&1
echo "== engine rc=$?"; tail -8 gpurun_out/pytest_engine.log
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -rA > gpurun_out/pytest_decoder.log 2>&1
echo "== decoder rc=$?"; tail -12 gpurun_out/pytest_decoder.log
cat gpurun_out/parity_report.json
