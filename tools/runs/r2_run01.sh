#!/bin/bash
# round 2, call 1: the new parity tests (wide fixtures, 256-line BASELINE workloads), the encoder-tail soak, the whole
# GPU suite and a bench line of the unchanged round-1 kernels (baseline for this round)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_01_newtests.log 2>&1; echo "== new tests rc=$?"; tail -15 gpurun_out/r2_01_newtests.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k soak > gpurun_out/r2_01_soak.log 2>&1; echo "== soak rc=$?"; tail -5 gpurun_out/r2_01_soak.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_wide_gpu.py --deselect tests/test_baseline_gpu.py > gpurun_out/r2_01_pytest_rest.log 2>&1; echo "== rest rc=$?"; tail -8 gpurun_out/r2_01_pytest_rest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_01_bench_fast.json 2> gpurun_out/r2_01_bench_fast.err; echo "== bench rc=$?"; cut -c1-400 gpurun_out/r2_01_bench_fast.json
cp gpurun_out/parity_report.json gpurun_out/r2_01_parity_report.json 2>/dev/null
