#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "== all rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke.log
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import torch, bench
from kiri_ocr_b200 import fixtures as FX
from kiri_ocr_b200.engine import BatchedRecognizer
cfg, tok, sd = bench.make_model()
for mode in ("bucketed",):
    eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode=mode)
    crops = FX.make_line_crops(256, seed=1234)
    buf, ent = eng.pack_crops(crops)
    for beam in (3, 5):
        cfg.BEAM = beam
        for _ in range(2): eng.recognize_packed(buf, ent, "beam")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): eng.recognize_packed(buf, ent, "beam")
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"beam={beam} {mode}: e2e {dt*1e3:.2f} ms per 256 lines = {256/dt:.0f} lines/s")
PY
