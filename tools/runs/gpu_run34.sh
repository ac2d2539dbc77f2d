#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_k.log 2>&1
echo "== kernels+engine rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_k.log | tail -8
KIRI_GEMM_TIMING=1 timeout 300 python tools/gemm_timing.py > gpurun_out/gemm_timing.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_timing.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), 'sync', round(d['e2e']['sync_value']), 'other', round(d['other_method']['value']))
print({k:(round(v['ms_per_step'],3)) for k,v in d['stages'].items()})
PY
