#!/bin/bash
# round 2, call 67: the beam-5 bench line on the final build
mkdir -p gpurun_out
timeout 60 python bench.py --method beam --steps 5 > gpurun_out/r02_final_bench_beam.json 2> gpurun_out/r02_final_bench_beam.err; echo "beam rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_final_bench_beam.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'steps',d['step_ms_min_p50_max'])
PY
