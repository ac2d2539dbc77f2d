#!/bin/bash
# round 2, call 24: preprocess: word-wise vertical pass (4 columns per lane), warp-per-row staging and padding
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -m gpu -q -x -k "prep or pillow or resize or crop or api or region" > gpurun_out/r2_24_k.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/r2_24_k.log
timeout 600 python tools/bench_hbm_kernels.py > gpurun_out/r2_24_hbm.json 2> gpurun_out/r2_24_hbm.err; echo rc=$?; grep -A3 preprocess gpurun_out/r2_24_hbm.json | grep -v "^--"
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'prep',round(d['stages']['preprocess']['ms_per_step'],4))"
