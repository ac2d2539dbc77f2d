#!/bin/bash
mkdir -p gpurun_out
for cs in 8 1; do
  KIRI_DEC_CLUSTER=$cs timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_api_gpu.py -m gpu -q -x > gpurun_out/pytest_dec_cs$cs.log 2>&1
  echo "== decoder tests cluster=$cs rc=$?"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_dec_cs$cs.log | tail -5
done
timeout 300 python tools/dec_timing.py > gpurun_out/dec_timing.log 2>&1; echo "rc=$?"; cat gpurun_out/dec_timing.log | tail -25
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'other',d['other_method'])
PY
