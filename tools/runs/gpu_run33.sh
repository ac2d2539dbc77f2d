#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',d['e2e'], 'other', round(d['other_method']['value']))
print({k:(round(v['ms_per_step'],3)) for k,v in d['stages'].items()})
PY
KIRI_BENCH_NO_SAMPLER=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_nosampler.json 2> gpurun_out/bench_nosampler.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_nosampler.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',d['e2e'], 'other', round(d['other_method']['value']))
PY
