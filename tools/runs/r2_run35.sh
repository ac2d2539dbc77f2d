#!/bin/bash
# round 2, call 35: stem sub-batch size sweep (does conv1's output stay in L2 for conv2 with smaller sub-batches?)
mkdir -p gpurun_out
for sc in 16 24 32 48 64 128; do
timeout 600 python bench.py --stem-chunk $sc 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
s=d['stages']
print('stem_chunk=$sc value',round(d['value']),'ms',round(d['ms_per_step'],4),'launches',d['gpu_launches'],'conv1',round(s['conv1']['ms_per_step'],4),'conv2',round(s['conv2']['ms_per_step'],4),'conv3',round(s['conv3']['ms_per_step'],4),'conv4',round(s['conv4']['ms_per_step'],4))"
done
