#!/bin/bash
# round 2, call 60: attention loads without the per-copy index decode (fixed swizzled chunk per thread, rows 48 apart);
# second leg: the same with a 104-register cap instead of launch bounds (96)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "attention or encoder_and_ctc or bucketed or multi_group" 2>&1 | tail -1
B='
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]
print(sys.argv[1],"value",round(d["value"]),"ms",round(d["ms_per_step"],4),"attention",round(d["stages"]["attention"]["ms_per_step"],4))'
timeout 300 python bench.py --steps 20 2>gpurun_out/r2_60_err.log | python -c "$B" launch_bounds
touch kiri-ocr_b200/csrc/attention.cu
make -C kiri-ocr_b200/csrc EXTRA="-DKIRI_ATTN_MAXNREG=104" > gpurun_out/r2_60_make.log 2>&1 || { echo make failed; exit 1; }
timeout 300 python bench.py --steps 20 2>>gpurun_out/r2_60_err.log | python -c "$B" maxnreg104
