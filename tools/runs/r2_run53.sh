#!/bin/bash
# round 2, call 53: conv2_swap with cp.async producers instead of 5-D TMA boxes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv" 2>&1 | tail -8
for v in cpasync tma; do
if [ $v = tma ]; then export KIRI_CONV2_TMA=1; else unset KIRI_CONV2_TMA; fi
timeout 300 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('$v fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'conv2',round(d['stages']['conv2']['ms_per_step'],4))"
done
