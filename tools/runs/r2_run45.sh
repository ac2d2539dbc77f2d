#!/bin/bash
# round 2, call 45: conv1 on the tensor pipe (one tcgen05 MMA pair per 128 pixels)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1" 2>&1 | tail -12
for v in tc ffma; do
if [ $v = ffma ]; then export KIRI_CONV1_FFMA=1; else unset KIRI_CONV1_FFMA; fi
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('$v fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'conv1',round(d['stages']['conv1']['ms_per_step'],4),'conv2',round(d['stages']['conv2']['ms_per_step'],4))"
done
unset KIRI_CONV1_FFMA
for c in 3 4; do
KIRI_CONV1_TC_CTAS=$c timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('ctas/sm=$c value',round(d['value']),'conv1',round(d['stages']['conv1']['ms_per_step'],4))"
done
