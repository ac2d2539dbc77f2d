"""A few launches of the fused encoder-layer tail at the bench's token count (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import eb_soak
lib = _lib.load(); _lib.require_device()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 26080
print("max|x|", eb_soak.soak(lib, M, 3))
