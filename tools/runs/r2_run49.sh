#!/bin/bash
# round 2, call 49: host side of submit() / collect(): plan_groups with one shared-memory pass, python scalars in collect
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_api_gpu.py tests/test_wide_gpu.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
e=d['e2e']
print('fast value',round(d['value']),'e2e',round(e['value']),'median submit/wait/collect',e['median_iter_ms_submit_wait_collect'],'phases',e['worst_iter_submit_phases_ms'])"
done
