#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dec_timing.py > gpurun_out/dec_timing.log 2>&1; echo "rc=$?"; cat gpurun_out/dec_timing.log | tail -30
