#!/bin/bash
# round 2, call 61: bf16 GEMM epilogue — chunk converted before the wait on the previous store's read, second chunk's
# TMEM read issued before the first chunk is staged (A/B against the serial form in the same box session)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "gemm or conv or encoder_and_ctc or bucketed or multi_group or encoder_block_fused" 2>&1 | tail -1
B='
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]
st=d["stages"]
print(sys.argv[1],"value",round(d["value"]),"ms",round(d["ms_per_step"],4)," ".join(k+"="+str(round(st[k]["ms_per_step"],4)) for k in ("qkv","conv3","conv4","pool_ln","ctc_head") if k in st))'
timeout 300 python bench.py --steps 20 2>gpurun_out/r2_61_err.log | python -c "$B" pipelined
touch kiri-ocr_b200/csrc/gemm_tc.cu
make -C kiri-ocr_b200/csrc EXTRA="-DKIRI_EPI_PIPE=0" > gpurun_out/r2_61_make.log 2>&1 || { echo make failed; exit 1; }
timeout 300 python bench.py --steps 20 2>>gpurun_out/r2_61_err.log | python -c "$B" serial
