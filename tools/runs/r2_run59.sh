#!/bin/bash
# round 2, call 59: QKV GEMM with two staging tiles per epilogue warp (and two A stages) instead of one (and four)
mkdir -p gpurun_out
touch kiri-ocr_b200/csrc/gemm_tc.cu
make -C kiri-ocr_b200/csrc EXTRA="-DKIRI_QKV_NBUF2=1" > gpurun_out/r2_59_make.log 2>&1 || { echo make failed; tail -5 gpurun_out/r2_59_make.log; exit 1; }
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('nbuf2 value',round(d['value']),'qkv',round(d['stages']['qkv']['ms_per_step'],4))"
