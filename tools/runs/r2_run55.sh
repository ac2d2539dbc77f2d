#!/bin/bash
# round 2, call 55: conv2_swap: the activation tile of a stage as 1 / 2 / 4 TMA boxes
mkdir -p gpurun_out
for sp in 2 4 1; do
touch kiri-ocr_b200/csrc/conv2_swap.cu
make -C kiri-ocr_b200/csrc EXTRA="-DKIRI_C2_SPLIT=$sp" > gpurun_out/r2_55_make.log 2>&1 || { echo make failed; tail -5 gpurun_out/r2_55_make.log; exit 1; }
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv2 or conv3x3" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('split=$sp value',round(d['value']),'conv2',round(d['stages']['conv2']['ms_per_step'],4))"
done
