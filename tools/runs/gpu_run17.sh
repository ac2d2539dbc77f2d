#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "== all rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke.log
for mode in bucketed parity; do
timeout 600 python bench.py --steps 20 --warmup 3 --width-mode $mode > gpurun_out/bench_fast_$mode.json 2> gpurun_out/bench_fast_$mode.err; echo "== bench $mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fast_$mode.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'other',d['other_method'])
print({k:(round(v['ms_per_step'],3), round(v.get('tflops',v.get('gbs',0)),1)) for k,v in d['stages'].items()})
PY
tail -3 gpurun_out/bench_fast_$mode.err
done
