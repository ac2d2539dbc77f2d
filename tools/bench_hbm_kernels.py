"""HBM-roofline exhibit of the two byte-bound kernels at a batch large enough to leave L2
(north_star: preprocessing and fused CTC against the measured HBM copy bandwidth)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kiri_ocr_b200 import fixtures as FX, _lib
from kiri_ocr_b200.engine import BatchedRecognizer

pk = bench.peaks()
cfg, tok, sd = bench.make_model()
eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode="parity")
lib = _lib.load()
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)

with torch.cuda.stream(eng.stream):
    for n_lines in (256, 2048, 8192):
        # ---- fused CTC greedy: [n, 160, 208] fp32 logits
        T, Cp, C = 160, eng.pw.Cp, eng.pw.C
        logits = torch.randn((n_lines, T, Cp), device="cuda")
        ids = torch.empty((n_lines, T), dtype=torch.int32, device="cuda")
        n_ids = torch.empty(n_lines, dtype=torch.int32, device="cuda"); conf = torch.empty(n_lines, device="cuda")
        ms = timeit(lambda: _lib.check(lib.kiri_ctc_greedy(logits.data_ptr(), 0, n_lines, T, C, Cp, ids.data_ptr(), n_ids.data_ptr(),
                                                           conf.data_ptr(), 0, 0, _lib.stream_ptr())))
        by = n_lines * (T * C * 4 + 4 * T + 12)
        out[f"ctc_greedy/{n_lines}"] = {"ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / pk["hbm_gbs"]}
        # ---- preprocess: the bench's synthetic crops repeated
        crops = FX.make_line_crops(256, seed=1234) * (n_lines // 256)
        buf, ent = eng.pack_crops(crops)
        prep = eng.prepare_resident(buf, ent)
        ms = timeit(lambda: _lib.check(lib.kiri_preprocess_pack(prep["src"].data_ptr(), prep["descs"].data_ptr(), prep["n_crops"], cfg.IMG_H,
                                                                prep["smem"], prep["n_strips"], prep["planes_all"].data_ptr(), 0, prep["sums"].data_ptr(), _lib.stream_ptr())))
        by = sum(c.size for c in crops) + n_lines * 48 * 640
        out[f"preprocess/{n_lines}"] = {"ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / pk["hbm_gbs"]}
        del prep, buf
print(json.dumps(out, indent=1))
