#!/bin/bash
# round 2, call 28: decode slots: clusters holding the longest lines get fewer lines (empty slots)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_28_t.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/r2_28_t.log
for n0 in 16 12 9 6 4; do
KIRI_DEC_SLOTS_N0=$n0 timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('n0=$n0 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
done
KIRI_DEC_TIMING=1 timeout 300 python tools/dec_timing.py 2>&1 | tail -45
