#!/bin/bash
# round 2, call 56: cProfile of the host side of submit() / collect()
mkdir -p gpurun_out
timeout 300 python tools/host_profile.py ctc > gpurun_out/r2_56_host_profile.txt 2>&1; tail -45 gpurun_out/r2_56_host_profile.txt | cut -c1-150
