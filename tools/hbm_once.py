"""One launch each of the two byte-bound kernels at N lines (for ncu byte counts): preprocess and the logits -> ids CTC kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kiri_ocr_b200 import fixtures as FX, _lib
from kiri_ocr_b200.engine import BatchedRecognizer
n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg, tok, sd = bench.make_model()
eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode="parity")
lib = _lib.load()
crops = FX.make_line_crops(256, seed=1234) * (n_lines // 256)
with torch.cuda.stream(eng.stream):
    buf, ent = eng.pack_crops(crops)
    prep = eng.prepare_resident(buf, ent)
    T, Cp, Cc = 160, eng.pw.Cp, eng.pw.C
    logits = torch.randn((n_lines, T, Cp), device="cuda")
    ids = torch.empty((n_lines, T), dtype=torch.int32, device="cuda")
    n_ids = torch.empty(n_lines, dtype=torch.int32, device="cuda"); conf = torch.empty(n_lines, device="cuda")
    for _ in range(2):
        _lib.check(lib.kiri_preprocess_pack(prep["src"].data_ptr(), prep["descs"].data_ptr(), prep["n_crops"], cfg.IMG_H, prep["smem"],
                                            prep["n_strips"], prep["planes_all"].data_ptr(), 0, prep["sums"].data_ptr(), _lib.stream_ptr()))
        _lib.check(lib.kiri_ctc_greedy(logits.data_ptr(), 0, n_lines, T, Cc, Cp, ids.data_ptr(), n_ids.data_ptr(), conf.data_ptr(), 0, 0,
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
alg_pre = sum(c.size for c in crops) + n_lines * 48 * 640
alg_ctc = n_lines * (T * Cc * 4 + 4 * T + 12)
print(f"algorithmic bytes: preprocess {alg_pre} ctc {alg_ctc}")
