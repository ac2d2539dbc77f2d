#!/bin/bash
# round 2, call 29: decoder: projection weights fetched ahead of the cluster barrier in front of each projection
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_beam_gpu.py tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_29_t.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/r2_29_t.log
for v in 0 1 0 1; do
if [ $v = 1 ]; then export KIRI_DEC_NO_WPRE=1; else unset KIRI_DEC_NO_WPRE; fi
timeout 600 python bench.py --method accurate 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('no_wpre=$v value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'dec_step',round(d['stages']['dec_step']['ms_per_step'],3))"
done
unset KIRI_DEC_NO_WPRE
timeout 600 python bench.py --method beam 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('beam value',round(d['value']),'ms',round(d['ms_per_step'],3))"
KIRI_DEC_TIMING=1 timeout 300 python tools/dec_timing.py 2>&1 | sed -n 1,26p
