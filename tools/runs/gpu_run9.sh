#!/bin/bash
# full GPU test suite + smoke + launch list of the accurate path (per-kernel durations of a decode step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rA > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -20
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 2 --warmup 3 --method accurate --width-mode parity > gpurun_out/plain_acc.log 2> gpurun_out/plain_acc.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 600 --csv --log-file gpurun_out/launches_acc.csv python bench.py --steps 2 --warmup 3 --method accurate --width-mode parity > gpurun_out/ncu_acc.log 2>&1
echo "== ncu rc=$?"
