#!/bin/bash
# round 2, call 15: conv2 skips the zero-filled K step (3 of 4 per tap); sub-phase timers of the encoder tail's E1 / E2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv or encoder_block" > gpurun_out/r2_15_k.log 2>&1; echo "== kernel tests rc=$?"; tail -5 gpurun_out/r2_15_k.log
timeout 200 python tools/eb_timing.py > gpurun_out/r2_15_eb_timing.txt 2>&1; echo rc=$?; cat gpurun_out/r2_15_eb_timing.txt
timeout 600 python bench.py > gpurun_out/r2_15_bench_fast.json 2> gpurun_out/r2_15_bench_fast.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_15_bench_fast.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_15_bench_fast.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3),'other',round(d['other_method']['value']))
print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
