#!/bin/bash
# round 2, call 12: CTC frame decisions in the head GEMM's epilogue (no logits in HBM for fast / accurate)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_12_pytest.log 2>&1; echo "== pytest rc=$?"; tail -25 gpurun_out/r2_12_pytest.log
timeout 600 python bench.py > gpurun_out/r2_12_bench_fast.json 2> gpurun_out/r2_12_bench_fast.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_12_bench_fast.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_12_bench_fast.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3),'other',round(d['other_method']['value']))
print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
