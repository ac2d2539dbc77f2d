"""Where does the host-facing (e2e) step go?  Per-iteration host timings of submit()/collect() and
device timings of the upload / compute, for the blocking and the two-in-flight forms.

    python tools/e2e_diag.py [--steps 30] [--width-mode bucketed]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--width-mode", default="bucketed")
    args = ap.parse_args()
    from kiri_ocr_b200 import fixtures as FX
    from kiri_ocr_b200.engine import BatchedRecognizer
    cfg, tok, sd = bench.make_model()
    eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode=args.width_mode)
    crops = FX.make_line_crops(256, seed=1234)
    buf, ent = eng.pack_crops(crops)
    for _ in range(3):
        eng.recognize_packed(buf, ent, "ctc")
    torch.cuda.synchronize()
    out = {}

    # raw H2D bandwidth of the source buffer (pinned)
    d = torch.empty(buf.numel(), dtype=torch.uint8, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        a.record(); d.copy_(buf, non_blocking=True); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out["h2d_ms"] = [round(t, 3) for t in ts]
    out["h2d_GBps"] = round(buf.numel() / (min(ts) / 1e3) / 1e9, 1)

    def run(mode):
        sub, col, tot = [], [], []
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        if mode == "sync":
            for _ in range(args.steps):
                t0 = time.perf_counter()
                tk = eng.submit(buf, ent, "ctc")
                t1 = time.perf_counter()
                eng.collect(tk)
                t2 = time.perf_counter()
                sub.append(t1 - t0); col.append(t2 - t1); tot.append(t2 - t0)
        else:
            t0 = time.perf_counter()
            tk = eng.submit(buf, ent, "ctc")
            sub.append(time.perf_counter() - t0)
            for _ in range(args.steps - 1):
                t0 = time.perf_counter()
                tk2 = eng.submit(buf, ent, "ctc")
                t1 = time.perf_counter()
                eng.collect(tk)
                t2 = time.perf_counter()
                sub.append(t1 - t0); col.append(t2 - t1); tot.append(t2 - t0)
                tk = tk2
            eng.collect(tk)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_all
        f = lambda v: [round(float(np.percentile(np.array(v) * 1e3, q)), 3) for q in (5, 50, 95, 100)]  # noqa: E731
        return {"lines_per_s": round(256 * args.steps / dt), "ms_per_step": round(dt / args.steps * 1e3, 3),
                "submit_ms_p5_50_95_max": f(sub), "collect_ms_p5_50_95_max": f(col), "iter_ms": f(tot)}

    for rep in range(2):
        out[f"sync_{rep}"] = run("sync")
        out[f"pipe_{rep}"] = run("pipe")
    # collect() split: wait vs host decode
    tk = eng.submit(buf, ent, "ctc")
    t0 = time.perf_counter(); tk["done"].synchronize(); t1 = time.perf_counter()
    eng.collect(tk); t2 = time.perf_counter()
    out["collect_after_done_ms"] = round((t2 - t1) * 1e3, 3)
    out["cpu_count"] = os.cpu_count()
    try:
        out["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
