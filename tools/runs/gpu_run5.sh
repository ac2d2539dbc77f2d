#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA -k "gemm or conv3x3" > gpurun_out/pytest_gemm.log 2>&1
echo "== gemm/conv rc=$?"; tail -6 gpurun_out/pytest_gemm.log
timeout 600 python bench.py --steps 20 --warmup 3 --width-mode parity > gpurun_out/bench_fast_parity.json 2> gpurun_out/bench_fast_parity.err; echo "== bench parity rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fast_parity.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])
print({k:(round(v['ms_per_step'],3), round(v.get('tflops',v.get('gbs',0)),1)) for k,v in d['stages'].items()})
PY
tail -3 gpurun_out/bench_fast_parity.err
