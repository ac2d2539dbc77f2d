#!/bin/bash
# round 2, call 11: decoder attention with lane-owned keys (dot products + private partial sums + butterfly)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_wide_gpu.py tests/test_beam_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/dec_timing.py > gpurun_out/r2_11_dec_timing.txt 2>&1; echo "rc=$?"; head -24 gpurun_out/r2_11_dec_timing.txt
timeout 600 python bench.py --method accurate --steps 10 > gpurun_out/r2_11_bench_acc.json 2> gpurun_out/r2_11_bench_acc.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_11_bench_acc.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_11_bench_acc.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3), 'dec_step ms', round(d['stages']['dec_step']['ms_per_step'],3))
PY
