"""Phase-level cycle breakdown (CTA 0) of the fused encoder-layer tail kernel (csrc/encoder_block.cu).
The sub-phase lines need a library built with `make -C kiri-ocr_b200/csrc EXTRA=-DKIRI_EB_SUBPHASE` (zeros otherwise); ptxas
moves the clock reads inside a basic block, so only sub-phases separated by a barrier wait or a branch are trustworthy."""
import ctypes as C, os, sys
os.environ["KIRI_GEMM_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib
lib = _lib.load(); _lib.require_device()
lib.kiri_debug_eb_timing.restype = C.c_int
lib.kiri_debug_eb_timing.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 32)()
NAMES = ["e1_wait_g1", "e1_wait_resid", "e1_work", "ff_wait_acc2_full", "ff_wait_h_empty", "ff_work", "e2_wait_x_full", "e2_work",
         "tiles", "mma_wait_ring", "mma_wait_a2", "mma_wait_h_full", "mma_wait_acc2_empty", "mma_total"]

def run(M, FF=1024, reps=10):
    D = 256
    dev = lambda t: t.cuda()
    o = dev((torch.randn(M, D) * 0.7).to(torch.bfloat16)); x = dev(torch.randn(M, D))
    wo = dev((torch.randn(D, D) / 16).to(torch.bfloat16)); w1 = dev((torch.randn(FF, D) / 16).to(torch.bfloat16))
    w2 = dev((torch.randn(D, FF) / 32).to(torch.bfloat16))
    bo, b1, b2 = dev(torch.zeros(D)), dev(torch.zeros(FF)), dev(torch.zeros(D))
    g = dev(torch.ones(D)); a = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
    fn = lambda: _lib.check(lib.kiri_encoder_block(o.data_ptr(), x.data_ptr(), a.data_ptr(), wo.data_ptr(), bo.data_ptr(), w1.data_ptr(),
                                                   b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), g.data_ptr(), bo.data_ptr(), g.data_ptr(),
                                                   bo.data_ptr(), M, FF, _lib.stream_ptr()))
    for _ in range(3): fn()
    torch.cuda.synchronize(); lib.kiri_debug_eb_timing(buf, 32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.kiri_debug_eb_timing(buf, 32)
    tiles = max(1, buf[8])
    fl = 2.0 * M * (256 * 256 + 2 * 256 * FF)
    print(f"encoder_block M={M} FF={FF}: {ms*1e3:7.1f} us/launch ({fl/ms/1e9:.0f} TF/s), CTA0 tiles/launch {tiles/reps:.1f}; cycles per tile: " +
          ", ".join(f"{n}={buf[i]/tiles:.0f}" for i, n in enumerate(NAMES) if n != "tiles"))
    print("   E1 sub-phases (wait residual | pass 1: +bo +resid -> X | row stats | pass 2: LN -> A2, +b2 -> X | st wait + fences + arrive): " +
          " ".join(f"{buf[16+i]/tiles:.0f}" for i in range(5)))
    print("   E2 sub-phases (read X + release + stage x + stores | row stats | LN in registers | wait x read | stage a + store | next residual): " +
          " ".join(f"{buf[23+i]/tiles:.0f}" for i in range(6)))

run(40960); run(26080); run(148 * 128)
