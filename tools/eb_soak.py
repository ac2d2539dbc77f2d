"""Soak of the fused encoder-layer tail (csrc/encoder_block.cu): `iters` back-to-back launches (PDL on) at several token
counts, twice from the same input; the two results must be bit-identical and finite.  With KIRI_GEMM_TIMING=1 in the
environment the kernel also runs its clock64 phase accounting (the configuration of the round-1 crash log).

    python tools/eb_soak.py [iters] [M ...]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib


def soak(lib, M, iters, FF=1024, with_ln=True, affine=True):
    D = 256
    g = torch.Generator().manual_seed(M)
    dev = lambda t: t.cuda()
    o = dev((torch.randn(M, D, generator=g) * 0.7).to(torch.bfloat16))
    x0 = dev(torch.randn(M, D, generator=g))
    wo = dev((torch.randn(D, D, generator=g) / 16).to(torch.bfloat16))
    w1 = dev((torch.randn(FF, D, generator=g) / 16).to(torch.bfloat16))
    w2 = dev((torch.randn(D, FF, generator=g) / 32).to(torch.bfloat16))
    bo, b1, b2 = dev(torch.randn(D, generator=g) * 0.1), dev(torch.randn(FF, generator=g) * 0.1), dev(torch.randn(D, generator=g) * 0.1)
    g1, h1 = dev(1 + 0.1 * torch.randn(D, generator=g)), dev(0.1 * torch.randn(D, generator=g))
    if not affine:                                      # identity affines select the parameter-free kernel variant (the engine's)
        g1, h1 = torch.ones_like(g1), torch.zeros_like(h1)
    outs = []
    for rep in range(2):
        x = x0.clone()
        a = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.kiri_encoder_block_soak(o.data_ptr(), x.data_ptr(), a.data_ptr() if with_ln else 0, wo.data_ptr(), bo.data_ptr(),
                                               w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), g1.data_ptr(), h1.data_ptr(),
                                               g1.data_ptr() if with_ln else 0, h1.data_ptr() if with_ln else 0, M, FF, iters,
                                               _lib.stream_ptr()), "kiri_encoder_block_soak")
        torch.cuda.synchronize()
        outs.append((x, a))
    assert torch.isfinite(outs[0][0]).all(), f"M={M}: non-finite x after {iters} launches"
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), f"M={M}: two soaks differ"
    return float(outs[0][0].abs().max())


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    Ms = [int(v) for v in sys.argv[2:]] or [128, 18944, 26080, 40960]
    lib = _lib.load()
    _lib.require_device()
    for M in Ms:
        for with_ln in (True, False):
            mx = soak(lib, M, iters, with_ln=with_ln, affine=with_ln)       # (affine variant with ln_out, folded variant without)
            print(f"soak ok: M={M} iters={iters} ln_out={with_ln} timing={'KIRI_GEMM_TIMING' in os.environ} "
                  f"lib={os.path.basename(_lib.LIB_PATH)} max|x|={mx:.1f}", flush=True)
