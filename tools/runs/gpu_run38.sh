#!/bin/bash
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_e2e_$i.json 2> gpurun_out/bench_e2e_$i.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_e2e_$i.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',d['e2e'])
PY
done
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6
