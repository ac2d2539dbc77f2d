#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "== bench_gemm rc=$?"; cat gpurun_out/bench_gemm.log
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -rA > gpurun_out/pytest_decoder.log 2>&1
echo "== decoder rc=$?"; tail -12 gpurun_out/pytest_decoder.log
timeout 600 python bench.py --steps 3 --warmup 3 --method accurate > gpurun_out/bench_acc.json 2> gpurun_out/bench_acc.err; echo "== bench acc rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_acc.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
PY
tail -3 gpurun_out/bench_acc.err
