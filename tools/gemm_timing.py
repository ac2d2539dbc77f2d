"""Role-level cycle breakdown (CTA 0) of the tcgen05 GEMM / conv kernel for the pipeline's shapes."""
import ctypes as C, os, sys
os.environ["KIRI_GEMM_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib
lib = _lib.load(); _lib.require_device()
lib.kiri_debug_gemm_timing.restype = C.c_int
lib.kiri_debug_gemm_timing.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 16)()
NAMES = ["tma_wait_empty", "tma_total", "mma_wait_full", "mma_wait_tmem_empty", "mma_total", "epi_wait_tmem_full",
         "epi_total", "epi_wait_resid", "epi_wait_store_read", "tiles"]

def report(tag, ms, reps):
    lib.kiri_debug_gemm_timing(buf, 16)
    tiles = max(1, buf[9])
    per = {n: buf[i] / tiles for i, n in enumerate(NAMES[:9])}
    print(f"{tag}: {ms*1e3:7.1f} us/launch, CTA0 tiles/launch {tiles/reps:.1f}; cycles per tile: " +
          ", ".join(f"{k}={v:.0f}" for k, v in per.items()))

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); lib.kiri_debug_gemm_timing(buf, 16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, reps

def gemm(M, N, K, epi):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    f32 = epi in (3, 4, 5)
    out = torch.zeros((M, N), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    out2 = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    g = torch.ones(N, device="cuda")
    fn = lambda: _lib.check(lib.kiri_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, N, K, epi, out.data_ptr(),
                                               out.data_ptr() if epi in (3, 5) else 0, g.data_ptr(), bias.data_ptr(), out2.data_ptr(), _lib.stream_ptr()))
    ms, reps = timeit(fn)
    report(f"gemm M={M} N={N} K={K} epi={epi}", ms, reps)

def conv(n, IH, IW, cin, cout, sh, sw, cin_mem=0):
    x = torch.randn(n, IH, IW, cin_mem or cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, 9 * cin, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    OH, OW = (IH + 2 - 3) // sh + 1, (IW + 2 - 3) // sw + 1
    out = torch.zeros((n, OH, OW, cout), dtype=torch.bfloat16, device="cuda")
    fn = lambda: _lib.check(lib.kiri_conv3x3_bf16(x.data_ptr(), w.data_ptr(), bias.data_ptr(), n, IH, IW, cin, cout, sh, sw,
                                                  out.data_ptr(), cin_mem, _lib.stream_ptr()))
    ms, reps = timeit(fn)
    fl = 2.0 * n * OH * OW * cout * 9 * cin
    report(f"conv n={n} {IH}x{IW}x{cin}->{cout} s({sh},{sw}) [{fl/ms/1e9:.0f} TF/s padded]", ms, reps)

gemm(26080, 768, 256, 0)
gemm(40960, 768, 256, 0)
gemm(40960, 256, 256, 5)
gemm(40960, 1024, 256, 2)
gemm(40960, 256, 1024, 5)
gemm(40960, 208, 256, 4)
conv(64, 48, 640, 64, 96, 2, 2, 48)
conv(64, 24, 320, 96, 160, 2, 2)
conv(64, 12, 160, 160, 256, 2, 1)

