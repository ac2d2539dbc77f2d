#!/bin/bash
# round 2, call 50: preprocess: both passes specialised on the tap count (3 / 5 / 7), no per-tap predicates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "prep or pillow or resize or crop or api or region or invert or engine" 2>&1 | tail -2
timeout 600 python tools/bench_hbm_kernels.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print({k:(round(v['ms'],4), round(v['frac_of_hbm_peak'],3)) for k,v in d.items() if 'prep' in k})"
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'prep',round(d['stages']['preprocess']['ms_per_step'],4))"
