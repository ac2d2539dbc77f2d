#!/bin/bash
# round 2, call 40: CTC logits kernel: 8 lanes per frame, 4 frames per warp pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_beam_gpu.py -m gpu -q -x -k "ctc or beam or engine" 2>&1 | tail -3
timeout 600 python tools/bench_hbm_kernels.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if 'ctc' in k: print(k, round(v['ms'],4), round(v['GBps']), round(v['frac_of_hbm_peak'],4))"
timeout 600 python bench.py --method beam 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('beam value',round(d['value']),'ms',round(d['ms_per_step'],3))"
