#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.json 2> gpurun_out/hbm_kernels.err; echo "rc=$?"; cat gpurun_out/hbm_kernels.json | tr -d '\n' | cut -c1-1500; tail -3 gpurun_out/hbm_kernels.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']), 'roof', d['roofline']['frac'], d['roofline']['traffic'])
print({k:(round(v['ms_per_step'],3), round(v.get('gbs',0)), round(v.get('frac_of_peak',0),3)) for k,v in d['stages'].items() if k in ('preprocess','ctc_greedy','conv1')})
PY
