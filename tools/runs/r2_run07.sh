#!/bin/bash
# round 2, call 7: GPU page ingest (BGR->gray), batched validation, live streaming (greedy + beam rule), decoder.cu rewrite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_07_pytest.log 2>&1; echo "== pytest rc=$?"; tail -25 gpurun_out/r2_07_pytest.log
timeout 600 python bench.py --method accurate --steps 5 > gpurun_out/r2_07_bench_acc.json 2> gpurun_out/r2_07_bench_acc.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_07_bench_acc.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_07_bench_acc.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3))
PY
