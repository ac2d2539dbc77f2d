#!/bin/bash
# round 2, call 34: encoder tail: the weights-only producer warp does not wait for the previous kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "encoder_block" 2>&1 | tail -2
timeout 200 python tools/eb_timing.py 2>&1 | grep -v sub-phases | cut -c1-330
for i in 1 2; do
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'tail',round(d['stages']['encoder_tail']['ms_per_step'],4))"
done
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_wide_gpu.py -m gpu -q -x 2>&1 | tail -2
