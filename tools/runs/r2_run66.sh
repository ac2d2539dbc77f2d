#!/bin/bash
# round 2, call 66: the pages workload with the final bench.py (heap frozen before the timed loop)
mkdir -p gpurun_out
timeout 85 python bench.py --workload pages --steps 3 > gpurun_out/r02_final_bench_pages.json 2> gpurun_out/r02_final_bench_pages.err; echo "pages rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_final_bench_pages.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'ordered_equal',d.get('config',{}).get('ordered_equal', d.get('ordered_equal')))
PY
