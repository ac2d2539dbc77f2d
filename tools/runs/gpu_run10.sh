#!/bin/bash
# fused decoder: parity at every cluster size, then accurate-mode timing
mkdir -p gpurun_out
for cs in 1 8 2 4; do
  KIRI_DEC_CLUSTER=$cs timeout 300 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > gpurun_out/pytest_dec_cs$cs.log 2>&1
  echo "== decoder tests cluster=$cs rc=$?"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_dec_cs$cs.log | tail -5
done
for cs in 8 4 1; do
KIRI_DEC_CLUSTER=$cs timeout 300 python bench.py --steps 5 --warmup 3 --method accurate --width-mode parity > gpurun_out/bench_acc_cs$cs.json 2> gpurun_out/bench_acc_cs$cs.err; echo "== bench acc cs=$cs rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_acc_cs$cs.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
PY
tail -3 gpurun_out/bench_acc_cs$cs.err
done
