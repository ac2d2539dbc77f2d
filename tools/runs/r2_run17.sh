#!/bin/bash
# round 2, call 17: A/B of the encoder tail on ONE box: the N=256 build of call 13 against the rewritten epilogues
mkdir -p gpurun_out
for v in r13 new r13 new; do
  cp tools/ab/encoder_block_$v.cu kiri-ocr_b200/csrc/encoder_block.cu
  make -C kiri-ocr_b200/csrc > gpurun_out/r2_17_make_$v.log 2>&1 || { echo "make $v failed"; tail -5 gpurun_out/r2_17_make_$v.log; exit 1; }
  echo "=== $v"
  timeout 200 python tools/eb_timing.py 2>&1 | grep -v "sub-phases" | cut -c1-420
  timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'tail',round(d['stages']['encoder_tail']['ms_per_step'],4),'conv2',round(d['stages']['conv2']['ms_per_step'],4),'clk',d['clocks']['sm_mhz'])"
done
