"""cProfile of the host side of submit() / collect() on the bench's 256-line batch (two batches in flight)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kiri_ocr_b200 import fixtures as FX
from kiri_ocr_b200.engine import BatchedRecognizer
cfg, tok, sd = bench.make_model()
eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode="bucketed")
crops = FX.make_line_crops(256, seed=1234)
buf, ent = eng.pack_crops(crops)
method = sys.argv[1] if len(sys.argv) > 1 else "ctc"
for _ in range(5):
    eng.collect(eng.submit(buf, ent, method))
pr = cProfile.Profile()
tk = eng.submit(buf, ent, method)
pr.enable()
for _ in range(300):
    tk2 = eng.submit(buf, ent, method)
    eng.collect(tk)
    tk = tk2
pr.disable()
eng.collect(tk)
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
