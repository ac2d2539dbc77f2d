#!/bin/bash
# round 2, call 31: decoder: cross K/V loads with an L2 evict-last policy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_beam_gpu.py -m gpu -q -x > gpurun_out/r2_31_t.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/r2_31_t.log
for v in 1 0 1 0; do
KIRI_DEC_KV_HINT=$v timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('hint=$v value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
done
KIRI_DEC_TIMING=1 timeout 300 python tools/dec_timing.py 2>&1 | sed -n 1,24p | grep -E "step_res|per step|F cross|B self"
timeout 600 python bench.py --method accurate --width-mode parity 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('parity hint=1 value',round(d['value']),'ms',round(d['ms_per_step'],3),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
KIRI_DEC_KV_HINT=0 timeout 600 python bench.py --method accurate --width-mode parity 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('parity hint=0 value',round(d['value']),'ms',round(d['ms_per_step'],3),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
