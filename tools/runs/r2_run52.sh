#!/bin/bash
# round 2, call 52: conv2 with swapped operands (channels as M, 256 pixels as N)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv" 2>&1 | tail -15
for v in swap noswap; do
if [ $v = noswap ]; then export KIRI_CONV2_NO_SWAP=1; else unset KIRI_CONV2_NO_SWAP; fi
timeout 300 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('$v fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'conv2',round(d['stages']['conv2']['ms_per_step'],4),'conv3',round(d['stages']['conv3']['ms_per_step'],4))"
done
