#!/bin/bash
mkdir -p gpurun_out
for sc in 8 16 24 32 64; do
timeout 300 python bench.py --steps 10 --warmup 3 --width-mode parity --stem-chunk $sc > gpurun_out/bench_sc$sc.json 2> gpurun_out/bench_sc$sc.err; echo "== stem_chunk $sc rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_sc$sc.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3))
print({k:round(v['ms_per_step'],3) for k,v in d['stages'].items() if k.startswith('conv')})
PY
done
