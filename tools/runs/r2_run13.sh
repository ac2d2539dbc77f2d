#!/bin/bash
# round 2, call 13: encoder tail with N = 256 MMAs throughout (hidden groups of 256), single-pass LayerNorm statistics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "encoder_block" > gpurun_out/r2_13_eb.log 2>&1; echo "== eb tests rc=$?"; tail -12 gpurun_out/r2_13_eb.log
timeout 200 python tools/eb_timing.py > gpurun_out/r2_13_eb_timing.txt 2>&1; echo rc=$?; cat gpurun_out/r2_13_eb_timing.txt
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_13_pytest.log 2>&1; echo "== pytest rc=$?"; tail -6 gpurun_out/r2_13_pytest.log
timeout 600 python bench.py > gpurun_out/r2_13_bench_fast.json 2> gpurun_out/r2_13_bench_fast.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_13_bench_fast.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_13_bench_fast.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3),'other',round(d['other_method']['value']))
print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
