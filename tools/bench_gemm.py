"""Micro-benchmark of the tcgen05 GEMM kernel: separates main-loop feed rate from per-tile overhead."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib
lib = _lib.load(); _lib.require_device()

def run(M, N, K, epi=_lib.EPI_BIAS_BF16, reps=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    f32 = epi in (_lib.EPI_BIAS_F32, _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_RESID_LN)
    out = torch.zeros((M, N), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    out2 = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    g = torch.ones(N, device="cuda")
    def call():
        _lib.check(lib.kiri_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, N, K, epi, out.data_ptr(),
                                      out.data_ptr() if epi in (3, 5) else 0, g.data_ptr(), bias.data_ptr(), out2.data_ptr(), _lib.stream_ptr()))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * N * K / ms / 1e9
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    per_cta = (tiles + 147) // 148
    cyc_tile = ms * 1e-3 * 1.9e9 / max(1, per_cta)
    print(f"M={M:7d} N={N:5d} K={K:5d} epi={epi}: {ms*1e3:8.1f} us  {tf:7.1f} TF/s  tiles/CTA={per_cta:3d}  cycles/tile~{cyc_tile:9.0f}  (MMA {K//16*128*min(N,256)//256})")

T = 128 * 148
for K in (64, 256, 1024, 4096):
    run(T, 256, K)
for K in (256, 1024, 4096):
    run(T * 8, 256, K)
for N in (64, 128, 256, 768, 1024):
    run(T * 4, N, 256)
for epi in (0, 1, 2, 3, 4, 5):
    run(T * 4, 256, 256, epi)
run(40960, 768, 256); run(40960, 1024, 256, 2); run(40960, 256, 1024, 5); run(40960, 256, 256, 5)
run(256, 768, 256); run(256, 256, 256, 5); run(256, 1024, 256, 2); run(256, 256, 1024, 5); run(256, 416, 256, 4)
