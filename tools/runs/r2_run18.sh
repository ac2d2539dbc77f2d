#!/bin/bash
# round 2, call 18: A/B/C of the encoder tail on ONE box: call-13 build | rewritten epilogues, run-time column quarter | compile-time quarter (4 x code)
mkdir -p gpurun_out
run() {  # $1 = source variant, $2 = extra flags, $3 = label
  cp tools/ab/encoder_block_$1.cu kiri-ocr_b200/csrc/encoder_block.cu
  make -C kiri-ocr_b200/csrc EXTRA="$2" > gpurun_out/r2_18_make_$3.log 2>&1 || { echo "make $3 failed"; tail -5 gpurun_out/r2_18_make_$3.log; exit 1; }
  echo "=== $3"
  timeout 200 python tools/eb_timing.py 2>&1 | grep "M=26080" | cut -c1-420
  timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'tail',round(d['stages']['encoder_tail']['ms_per_step'],4),'clk',d['clocks']['sm_mhz'])"
}
run r13 "" r13
run new "" new_runtime_cq
run new "-DKIRI_EB_CQ_SWITCH" new_switch
run new "-DKIRI_EB_SUBPHASE" new_runtime_cq_subphase
timeout 200 python tools/eb_timing.py 2>&1 | grep -A2 "M=26080" | cut -c1-420
