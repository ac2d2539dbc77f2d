#!/bin/bash
# round 2, call 62: pool kernel with its six row loads in flight at once; final LayerNorm without the bf16 mem store on the fast path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_api_gpu.py -m gpu -q -x -k "pool or encoder_and_ctc or bucketed or multi_group or goldens or extract_text or pages" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 2>gpurun_out/r2_62_err.log | python -c '
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]
st=d["stages"]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"])," ".join(k+"="+str(round(st[k]["ms_per_step"],4)) for k in ("pool_ln","ln_final","ctc_head","qkv","attention") if k in st))'
