#!/bin/bash
# round 2, call 41: CTC logits kernel: occupancy / loads-in-flight variants
mkdir -p gpurun_out
for cfg in "1 4" "2 3" "2 2" "1 6"; do set -- $cfg
touch kiri-ocr_b200/csrc/ctc.cu
make -C kiri-ocr_b200/csrc EXTRA="-DKIRI_CTC_IT=$1 -DKIRI_CTC_MINB=$2" > gpurun_out/r2_41_make.log 2>&1 || { echo make failed; tail -5 gpurun_out/r2_41_make.log; exit 1; }
timeout 600 python tools/bench_hbm_kernels.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('IT=$1 MINB=$2', {k:(round(v['ms'],4), round(v['frac_of_hbm_peak'],3)) for k,v in d.items() if 'ctc' in k})"
done
