#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA > gpurun_out/pytest_engine.log 2>&1
echo "== engine rc=$?"; tail -8 gpurun_out/pytest_engine.log
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -rA > gpurun_out/pytest_decoder.log 2>&1
echo "== decoder rc=$?"; tail -12 gpurun_out/pytest_decoder.log
cat gpurun_out/parity_report.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; echo "== bench rc=$?"; tail -c 3000 gpurun_out/bench_fast.json; tail -5 gpurun_out/bench_fast.err
timeout 600 python bench.py --steps 10 --warmup 3 --width-mode parity > gpurun_out/bench_fast_parity.json 2> gpurun_out/bench_fast_parity.err; echo "== bench parity rc=$?"; tail -c 1500 gpurun_out/bench_fast_parity.json
timeout 600 python bench.py --steps 3 --warmup 3 --method accurate > gpurun_out/bench_acc.json 2> gpurun_out/bench_acc.err; echo "== bench acc rc=$?"; tail -c 1500 gpurun_out/bench_acc.json; tail -5 gpurun_out/bench_acc.err
