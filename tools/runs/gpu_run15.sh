#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --width-mode parity > gpurun_out/plain_fast.log 2> gpurun_out/plain_fast.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 180 --csv --log-file gpurun_out/launches_fast_v2.csv python bench.py --steps 2 --warmup 3 --width-mode parity > gpurun_out/ncu_fast.log 2>&1
echo "== ncu rc=$?"
