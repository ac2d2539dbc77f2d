#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv3x3" > gpurun_out/pytest_gemm.log 2>&1
echo "== gemm/conv rc=$?"; grep -E "passed|failed|FAILED|Error|timeout" gpurun_out/pytest_gemm.log | tail -12
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "== bench_gemm rc=$?"; cat gpurun_out/bench_gemm.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "== all rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -8
for mode in parity bucketed; do
timeout 600 python bench.py --steps 20 --warmup 3 --width-mode $mode > gpurun_out/bench_fast_$mode.json 2> gpurun_out/bench_fast_$mode.err; echo "== bench $mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fast_$mode.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])
print({k:(round(v['ms_per_step'],3), round(v.get('tflops',v.get('gbs',0)),1)) for k,v in d['stages'].items()})
PY
tail -3 gpurun_out/bench_fast_$mode.err
done
