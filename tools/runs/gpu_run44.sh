#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_beam_gpu.py tests/test_api_gpu.py -m gpu -q > gpurun_out/pytest_dec.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/pytest_dec.log
timeout 600 python bench.py --steps 20 --warmup 3 --method accurate > gpurun_out/bench_acc.json 2> gpurun_out/bench_acc.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_acc.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))
print({k:(round(v['ms_per_step'],3)) for k,v in d['stages'].items() if k.startswith('dec')})
PY
