#!/bin/bash
# round 2, call 14: ncu full-set capture (with source) of the fused encoder tail
mkdir -p gpurun_out
python tools/eb_once.py > gpurun_out/r2_14_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:encoder_block -s 2 -c 1 -f -o gpurun_out/r2_14_eb python tools/eb_once.py > gpurun_out/r2_14_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_14_ncu.log; ls -la gpurun_out/r2_14_eb.ncu-rep
