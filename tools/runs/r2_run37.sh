#!/bin/bash
# round 2, call 37: soak of the encoder tail incl. the KIRI_CHECKED build (device-side invariant checks)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "soak" 2>&1 | tail -3
KIRI_B200_LIB=kiri-ocr_b200/libkiri_b200_checked.so timeout 600 python tools/eb_soak.py 50 26080 2>&1 | tail -3
KIRI_B200_LIB=$PWD/kiri-ocr_b200/libkiri_b200_checked.so timeout 600 python bench.py --steps 5 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('checked build: fast value',round(d['value']),'ms',round(d['ms_per_step'],4))"
