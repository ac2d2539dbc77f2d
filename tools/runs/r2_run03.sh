#!/bin/bash
# round 2, call 3: engine rework (decoder enqueued in submit, persistent staging, multi-page path, no-invert flag, isolation),
# new bench modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_03_pytest.log 2>&1; echo "== pytest rc=$?"; tail -12 gpurun_out/r2_03_pytest.log
for m in fast accurate beam; do
timeout 600 python bench.py --method $m > gpurun_out/r2_03_bench_$m.json 2> gpurun_out/r2_03_bench_$m.err; echo "== bench $m rc=$?"; tail -3 gpurun_out/r2_03_bench_$m.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_03_bench_$m.json'))
    print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['sync_value']),'launches',d['gpu_launches'],'roof',d['roofline']['kernel'][:28],round(d['roofline']['frac'],3),'other',d['other_method'] and round(d['other_method']['value']))
except Exception as e: print('parse failed',e)
PY
done
timeout 900 python bench.py --workload pages > gpurun_out/r2_03_bench_pages.json 2> gpurun_out/r2_03_bench_pages.err; echo "== bench pages rc=$?"; tail -3 gpurun_out/r2_03_bench_pages.err; cut -c1-250 gpurun_out/r2_03_bench_pages.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_03_bench_pages.json'))
    print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'equal',d['ordered_equal_to_single_gpu'],'launches',d['gpu_launches'])
except Exception as e: print('parse failed',e)
PY
