#!/bin/bash
# round 2, call 64: bench lines with the heap frozen before the device-timed loop (accurate, then fast)
O=gpurun_out/r02_final2
mkdir -p gpurun_out
S='
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]),"steps",d["step_ms_min_p50_max"],"launches",d["gpu_launches"],"roof",round(d["roofline"]["frac"],3))'
timeout 300 python bench.py --method accurate > ${O}_bench_accurate.json 2> ${O}_bench_accurate.err; echo "accurate rc=$?"; python -c "$S" ${O}_bench_accurate.json
timeout 300 python bench.py > ${O}_bench_fast.json 2> ${O}_bench_fast.err; echo "fast rc=$?"; python -c "$S" ${O}_bench_fast.json
