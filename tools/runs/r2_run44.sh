#!/bin/bash
# round 2, call 44: conv1: 2 / 4 / 8 image rows per thread (one constant-bank weight fetch feeds ROWS FFMA2s)
mkdir -p gpurun_out
for r in 2 4 8; do
export KIRI_CONV1_ROWS=$r
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv1" 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('rows=$r fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'conv1',round(d['stages']['conv1']['ms_per_step'],4))"
done
