"""One conv3-shaped launch (KC = 32, three chunks per k-block) — for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib
lib = _lib.load(); _lib.require_device()
n, IH, IW, cin, cout, sh, sw = 1, 24, 64, 96, 160, 2, 2
x = torch.randn(n, IH, IW, cin, device="cuda").to(torch.bfloat16)
w = torch.randn(cout, 9 * cin, device="cuda").to(torch.bfloat16)
bias = torch.zeros(cout, device="cuda")
OH, OW = (IH + 2 - 3) // sh + 1, (IW + 2 - 3) // sw + 1
out = torch.zeros((n, OH, OW, cout), dtype=torch.bfloat16, device="cuda")
_lib.check(lib.kiri_conv3x3_bf16(x.data_ptr(), w.data_ptr(), bias.data_ptr(), n, IH, IW, cin, cout, sh, sw, out.data_ptr(), 0, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
