"""A few launches of the preprocess kernels on the bench's synthetic crops (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kiri_ocr_b200 import fixtures as FX, _lib
from kiri_ocr_b200.engine import BatchedRecognizer
n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg, tok, sd = bench.make_model()
eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode="bucketed")
lib = _lib.load()
crops = FX.make_line_crops(256, seed=1234) * (n_lines // 256)
with torch.cuda.stream(eng.stream):
    buf, ent = eng.pack_crops(crops)
    prep = eng.prepare_resident(buf, ent)
    for _ in range(3):
        _lib.check(lib.kiri_preprocess_pack(prep["src"].data_ptr(), prep["descs"].data_ptr(), prep["n_crops"], cfg.IMG_H, prep["smem"],
                                            prep["n_strips"], prep["planes_all"].data_ptr(), 0, prep["sums"].data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
print("ok", prep["n_crops"], prep["n_strips"], prep["smem"])
