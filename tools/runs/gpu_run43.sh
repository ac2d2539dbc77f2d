#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "ctc or engine or recognize or goldens" > gpurun_out/pytest_ctc.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/pytest_ctc.log
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.json 2> gpurun_out/hbm_kernels.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/hbm_kernels.json'))
for k,v in d.items(): print(k, round(v['ms'],4),'ms', round(v['GBps']),'GB/s', round(v['frac_of_hbm_peak'],3))
PY
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))
print({k:(round(v['ms_per_step'],3)) for k,v in d['stages'].items()})
PY
