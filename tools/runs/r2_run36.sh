#!/bin/bash
# round 2, call 36: decoder attention: the first V row of a group fetched together with the first K row
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_beam_gpu.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('accurate value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
done
