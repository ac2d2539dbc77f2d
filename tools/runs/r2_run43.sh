#!/bin/bash
# round 2, call 43: ncu source-level capture of the current preprocess_pack_kernel (2048 lines)
mkdir -p gpurun_out
timeout 300 python tools/prep_once.py 2048 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:preprocess_pack -s 2 -c 1 -o gpurun_out/r2_43_prep -f python tools/prep_once.py 2048 > gpurun_out/r2_43_ncu.log 2>&1; echo "ncu rc=$?"
