#!/bin/bash
mkdir -p gpurun_out
for mode in parity bucketed; do
timeout 600 python bench.py --steps 10 --warmup 3 --method accurate --width-mode $mode > gpurun_out/bench_acc_$mode.json 2> gpurun_out/bench_acc_$mode.err; echo "== bench acc $mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_acc_$mode.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['stages'].items()})
PY
done
