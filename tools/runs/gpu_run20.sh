#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "== all rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -8
timeout 300 python tools/dec_timing.py > gpurun_out/dec_timing.log 2>&1; echo "rc=$?"; tail -18 gpurun_out/dec_timing.log
for mode in parity bucketed; do
timeout 600 python bench.py --steps 10 --warmup 3 --method accurate --width-mode $mode > gpurun_out/bench_acc_$mode.json 2> gpurun_out/bench_acc_$mode.err; echo "== bench acc $mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_acc_$mode.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['stages'].items() if k.startswith('dec')})
PY
done
