"""Per-phase cycle breakdown of the fused decode kernel (cluster 0 / CTA 0), KIRI_DEC_TIMING=1."""
import ctypes as C, os, sys, time
os.environ["KIRI_DEC_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kiri_ocr_b200 import fixtures as FX, _lib
from kiri_ocr_b200.engine import BatchedRecognizer
cfg, tok, sd = bench.make_model()
eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode="parity")
crops = FX.make_line_crops(256, seed=1234)
buf, ent = eng.pack_crops(crops)
prep = eng.prepare_resident(buf, ent)
lib = _lib.load()
lib.kiri_debug_decode_timing.restype = C.c_int
lib.kiri_debug_decode_timing.argtypes = [C.POINTER(C.c_longlong), C.c_int]
names = ["S0 embed+LN", "A qkv", "B self-attn", "B csync", "C out-proj", "C csync", "D LN2", "E cross-q", "F cross-attn",
         "F csync", "G cross-out", "G csync", "H LN3", "I ff1", "I csync", "J ff2", "J csync", "K LN", "heads", "heads csync", "select"]
with torch.cuda.stream(eng.stream):
    for it in range(3):
        out = eng.step_resident(prep, "decoder")
        torch.cuda.synchronize()
        buf_ = (C.c_longlong * 32)()
        lib.kiri_debug_decode_timing(buf_, 32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.step_resident(prep, "decoder"); e1.record(); torch.cuda.synchronize()
    lib.kiri_debug_decode_timing(buf_, 32)
n_out = out[0][1].cpu().numpy()
steps = int(n_out.max())
tot = sum(buf_[:21])
print(f"step_resident(decoder) {e0.elapsed_time(e1):.3f} ms; max steps {steps}; mean steps {n_out.mean():.1f}; cluster0 steps {int(n_out[:16].max())}")
print(f"cluster 0 total {tot} cycles = {tot/1.9e3:.1f} us; per step {tot/max(1,int(n_out[:16].max()))/1.9e3:.1f} us")
for i, nme in enumerate(names):
    print(f"  {nme:14s} {buf_[i]:10d} cyc  {100*buf_[i]/tot:5.1f}%  per step {buf_[i]/max(1,int(n_out[:16].max())):8.0f}")

lib.kiri_debug_decode_clusters.restype = C.c_int
lib.kiri_debug_decode_clusters.argtypes = [C.POINTER(C.c_longlong), C.c_int]
cb = (C.c_longlong * 256)()
lib.kiri_debug_decode_clusters(cb, 256)
t0 = min(cb[c * 4] for c in range(16))
print("cluster: start_us end_us dur_us steps smid  (lines' max steps)")
for c in range(16):
    print(f"  {c:2d}: {(cb[c*4]-t0)/1e3:8.1f} {(cb[c*4+1]-t0)/1e3:8.1f} {(cb[c*4+1]-cb[c*4])/1e3:8.1f} {cb[c*4+2]:4d} {cb[c*4+3]:4d}   {int(n_out[c*16:(c+1)*16].max())}")
