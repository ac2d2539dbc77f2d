#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rA > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -20
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke.log
