#!/bin/bash
# round 2, call 33: pages workload: batch size sweep (per-batch host overhead vs latency)
mkdir -p gpurun_out
for bl in 320 640 960 1280; do
timeout 600 python bench.py --workload pages --batch-lines $bl 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('batch_lines=$bl value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
done
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('fast value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"
timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -2
