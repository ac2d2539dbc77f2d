"""On-device diagnosis of the tcgen05 GEMM: isolates rows / K-chunks / columns on failure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kiri_ocr_b200 import _lib

lib = _lib.load(); _lib.require_device()

def run(M, N, K, a, w):
    bias = torch.zeros(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.kiri_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, N, K, _lib.EPI_BIAS_F32, out.data_ptr(), 0, 0, 0, 0, _lib.stream_ptr())
    if rc: print("rc", rc, lib.kiri_last_error()); return None
    torch.cuda.synchronize()
    return out

def case(M, N, K, tag=""):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ref = a.float() @ w.float().t()
    out = run(M, N, K, a, w)
    if out is None: return False
    err = (out - ref).abs()
    nan = torch.isnan(out).sum().item()
    print(f"[{tag}] M={M} N={N} K={K}: maxerr={err[~torch.isnan(err)].max().item() if nan < out.numel() else float('nan'):.4g} nan={nan}")
    ok = nan == 0 and err.max().item() < 0.05
    if not ok:
        bad = (err > 0.05) | torch.isnan(out)
        rows = bad.any(1).nonzero().flatten().tolist(); cols = bad.any(0).nonzero().flatten().tolist()
        print("  bad rows (first 40):", rows[:40], "count", len(rows))
        print("  bad cols (first 40):", cols[:40], "count", len(cols))
        # K-chunk isolation: only chunk c non-zero
        for c in range(K // 32):
            a2 = torch.zeros_like(a); a2[:, c*32:(c+1)*32] = a[:, c*32:(c+1)*32]
            o2 = run(M, N, K, a2, w); r2 = a2.float() @ w.float().t()
            e2 = (o2 - r2).abs(); print(f"   chunk {c}: maxerr {e2.max().item():.4g}")
        # does the output look like a permutation of rows? compare row 1 of out to all ref rows
        for r in (0, 1, 8, 9):
            if r < M:
                d = (ref - out[r:r+1]).abs().max(1).values
                print(f"   out row {r} best matches ref row {int(d.argmin())} (err {d.min().item():.3g})")
    return ok

ok = True
for (M, N, K) in [(128, 16, 64), (128, 64, 64), (128, 256, 64), (128, 256, 256), (256, 256, 256), (200, 208, 256), (512, 768, 256), (4096, 256, 1024)]:
    ok &= case(M, N, K, "gemm")
print("DIAG", "OK" if ok else "FAIL")
