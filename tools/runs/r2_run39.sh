#!/bin/bash
# round 2, call 39: preprocess: the strips of a crop are one cluster that also takes the crop sum (no crop_sum_kernel launch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "prep or pillow or resize or crop or api or region or invert or engine" 2>&1 | tail -3
timeout 600 python tools/bench_hbm_kernels.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if 'preprocess' in k: print(k, round(v['ms'],4), round(v['GBps']), round(v['frac_of_hbm_peak'],4))"
for v in 0 1; do
if [ $v = 1 ]; then export KIRI_PRE_NO_CLUSTER=1; else unset KIRI_PRE_NO_CLUSTER; fi
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('no_cluster=$v fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'prep',round(d['stages']['preprocess']['ms_per_step'],4))"
done
