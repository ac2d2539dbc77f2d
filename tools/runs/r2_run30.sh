#!/bin/bash
# round 2, call 30: decoder cross-attention: a line's keys split over idle warps (clusters with <= 8 live lines)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_beam_gpu.py tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_30_t.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/r2_30_t.log
for n0 in 9 8 6 5 4 3 2; do
KIRI_DEC_SLOTS_N0=$n0 timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('n0=$n0 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
done
KIRI_DEC_NO_KSPLIT=1 KIRI_DEC_SLOTS_N0=4 timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('nosplit n0=4 value',round(d['value']),'ms',round(d['ms_per_step'],3),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
KIRI_DEC_SLOTS_N0=4 KIRI_DEC_TIMING=1 timeout 300 python tools/dec_timing.py 2>&1 | sed -n 1,30p
