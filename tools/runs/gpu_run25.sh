#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "== all rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -8
for nopdl in 0 1; do
if [ $nopdl = 1 ]; then export KIRI_NO_PDL=1; fi
for mode in bucketed parity; do
timeout 600 python bench.py --steps 20 --warmup 3 --width-mode $mode > gpurun_out/bench_pdl${nopdl}_$mode.json 2> gpurun_out/bench_pdl${nopdl}_$mode.err; echo "== bench nopdl=$nopdl $mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl${nopdl}_$mode.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'other',round(d['other_method']['value']))
PY
done
done
