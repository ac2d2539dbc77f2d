#!/bin/bash
# round 2, call 68: the text-equality tests after the host-side decode change (whatever fits in the seconds left)
timeout 19 python -m pytest tests/test_wide_gpu.py -m gpu -q -x -k "fast_equals or accurate_equals" 2>&1 | tail -1
