#!/bin/bash
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null
for i in 1 2 3 4; do
timeout 600 python bench.py --steps 60 --warmup 3 > gpurun_out/bench_e2e_$i.json 2> gpurun_out/bench_e2e_$i.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_e2e_$i.json'))
e=d['e2e']
print('value',round(d['value']),'e2e',round(e['value']), e['iter_ms_p50_p95_max'], e['worst_iter_ms_submit_wait_collect'], e['worst_iter_submit_phases_ms'])
PY
done
