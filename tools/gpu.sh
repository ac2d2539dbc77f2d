#!/bin/bash
# build the library, then run a script on the GPU box:  tools/gpu.sh <timeout s> <script> [gpus]
set -e
make -C "$(dirname "$0")/../kiri-ocr_b200/csrc" -j8 all checked > /tmp/kiri_make.log 2>&1 || { tail -30 /tmp/kiri_make.log; exit 1; }
G=""; [ -n "$3" ] && G="--gpus $3"
exec gpurun $G --timeout "$1" -- "bash $2"
