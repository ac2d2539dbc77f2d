#!/bin/bash
# round 2, call 57: stream handle pinned inside submit() / collect()
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_api_gpu.py tests/test_wide_gpu.py tests/test_decoder_gpu.py tests/test_beam_gpu.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
e=d['e2e']
print('fast value',round(d['value']),'e2e',round(e['value']),'median submit/wait/collect',e['median_iter_ms_submit_wait_collect'])"
done
