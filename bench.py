"""bench.py — text lines/sec of the batched line-recognition hot path on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3               # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # reference algorithm on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (preprocess -> stem -> encoder -> CTC head -> CTC greedy)
over one batch of 256 synthetic line crops per GPU — BASELINE.json configs[1].  ``value`` is
timed with CUDA events with the packed source crops already in HBM; ``e2e`` goes through the
public call (``BatchedRecognizer.recognize_packed``) from pinned host memory to Python strings.
Every rank works on its own crops (weak scaling); the only collective is one NCCL all-gather of
the fixed-stride result records per step.  One JSON line is printed by rank 0.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "text lines/sec (CTC fast + accurate decode) at 1/2/4/8 B200; encoder tensor-pipe % peak"
BATCH = 256


def flops_per_line(Wb: int, C: int = 204) -> float:
    """2*(1 486 080*Wb + 4*(786 432*T + 512*T^2) + 256*C*T), T = Wb/4."""
    T = Wb // 4
    return 2.0 * (1486080 * Wb + 4 * (786432 * T + 512 * T * T) + 256 * C * T)


STAGE_FLOPS_PER_LINE = {           # algorithmic (unpadded) FLOPs of one line at width Wb
    "conv2": lambda Wb: 2.0 * (24 * (Wb // 2)) * 96 * 9 * 48,
    "conv3": lambda Wb: 2.0 * (12 * (Wb // 4)) * 160 * 9 * 96,
    "conv4": lambda Wb: 2.0 * (6 * (Wb // 4)) * 256 * 9 * 160,
    "qkv": lambda Wb: 4 * 2.0 * (Wb // 4) * 768 * 256,
    "out_proj": lambda Wb: 4 * 2.0 * (Wb // 4) * 256 * 256,
    "ff1": lambda Wb: 4 * 2.0 * (Wb // 4) * 1024 * 256,
    "ff2": lambda Wb: 4 * 2.0 * (Wb // 4) * 1024 * 256,
    # out_proj + linear1 + linear2 of all four layers in the fused kernel (csrc/encoder_block.cu)
    "encoder_tail": lambda Wb: 4 * 2.0 * (Wb // 4) * (256 * 256 + 2 * 1024 * 256),
    "attention": lambda Wb: 4 * 2.0 * 2 * 256 * (Wb // 4) ** 2,
    "ctc_head": lambda Wb: 2.0 * (Wb // 4) * 204 * 256,
}
STAGE_BYTES_PER_LINE = {           # algorithmic bytes for the CUDA-core / HBM-bound stages
    "conv1": lambda Wb: 48 * Wb + 48 * Wb * 48 * 2,
    "ctc_greedy": lambda Wb: (Wb // 4) * 204 * 4 + 4 * (Wb // 4) + 12,      # fp32 logits in, ids + stats out
}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p.update(json.load(open(path)))
            p["src"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        lo, hi = (self.t0 or 0.0) - 0.03, (self.t1 or 1e18) + 0.03
        rows = [r for t, r in self.rows if lo <= t <= hi] or [r for t, r in self.rows[-3:]]
        self.rows = rows
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_model():
    from kiri_ocr_b200 import fixtures as FX
    from kiri_ocr_b200.config import CFG, CharTokenizer
    cfg = CFG()
    d = tempfile.mkdtemp(prefix="kiri_bench_")
    vp = os.path.join(d, "vocab.json")
    with open(vp, "w", encoding="utf-8") as f:
        json.dump(FX.make_vocab(), f, ensure_ascii=False)
    tok = CharTokenizer(vp, cfg)
    sd = FX.make_state_dict(cfg, tok.vocab_size, seed=0, hardened=False)     # random-init weights
    return cfg, tok, sd


def cpu_baseline(cfg, tok, sd, crops, method: str, budget_s: float, min_lines: int = 8):
    """The oracle port of the reference algorithm, one line at a time (the reference's only mode,
    core.py:770-776), fp32, all host threads."""
    from oracle import decode as OD, preprocess as OP
    torch.set_num_threads(os.cpu_count() or 1)
    OD.recognize_plane(sd, tok, cfg, OP.preprocess_crop(crops[0]), method)           # warm-up
    t0 = time.perf_counter()
    n = 0
    while n < len(crops) and (n < min_lines or time.perf_counter() - t0 < budget_s):
        OD.recognize_plane(sd, tok, cfg, OP.preprocess_crop(crops[n]), method)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kiri_ocr_b200 import fixtures as FX
    cfg, tok, sd = make_model()
    crops = FX.make_line_crops(BATCH, seed=1234)
    method = "ctc" if args.method == "fast" else "decoder"
    per_step = 16 if method == "ctc" else 4
    from oracle import decode as OD, preprocess as OP
    torch.set_num_threads(os.cpu_count() or 1)
    k = 0

    def step():
        nonlocal k
        for _ in range(per_step):
            c = crops[k % len(crops)]
            k += 1
            OD.recognize_plane(sd, tok, cfg, OP.preprocess_crop(c), method)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f"{per_step} lines/step of the same 256-crop workload, one line at a time (reference mode), fp32, "
              f"oracle port with KV-cached decoder and a single encode per line")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "lines/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "decode_method": args.method, "lines_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    return (f"batch {BATCH} synthetic line crops per GPU (heights 24-96 px, widths bucketed to "
            f"{{128,256,384,512,640}} at H=48), random-init kiri recognizer V=202, decode_method={args.method}, "
            f"width_mode={args.width_mode}")


def run_ours(args):
    import torch.distributed as dist
    from kiri_ocr_b200 import _lib, fixtures as FX
    from kiri_ocr_b200.engine import BatchedRecognizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("KIRI_BENCH_NO_SAMPLER"):
        sampler.start()                                     # nvidia-smi needs a while to come up
    cfg, tok, sd = make_model()
    eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode=args.width_mode, stem_chunk=args.stem_chunk)
    method = "ctc" if args.method == "fast" else "decoder"
    crops = FX.make_line_crops(BATCH, seed=1234 + rank)
    buf, ent = eng.pack_crops(crops)
    prep = eng.prepare_resident(buf, ent)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    T = cfg.IMG_W // 4
    rec_w = 2 + T
    gathered = torch.empty((world * BATCH, rec_w), dtype=torch.int32, device="cuda") if world > 1 else None
    rec_buf = torch.zeros((BATCH, rec_w), dtype=torch.int32, device="cuda")

    def gather(outs):
        """The path's one exchange step: fixed-stride records {n, conf bits, ids[T]} to every rank."""
        if world == 1:
            return
        if len(outs) == 1 and outs[0][0].dim() == 1:
            # CTC: one kernel turns the token-major ids into the fixed-stride records
            ids, n, conf = outs[0]
            _lib.check(eng.lib.kiri_pack_records(ids.data_ptr(), n.data_ptr(), conf.data_ptr(), prep["mem_row0"].data_ptr(),
                                                 BATCH, T, rec_buf.data_ptr(), _lib.stream_ptr()), "kiri_pack_records")
            eng.launches += 1
            dist.all_gather_into_tensor(gathered, rec_buf)
            return
        rec = torch.zeros((BATCH, rec_w), dtype=torch.int32, device="cuda")
        r0 = 0
        for o in outs:
            ids, n, conf = o[0], o[1], o[-1]
            k = n.shape[0]
            rec[r0:r0 + k, 0] = n
            rec[r0:r0 + k, 1] = conf.view(torch.int32)
            if ids.dim() == 1:
                continue                                     # CTC: packed by kiri_pack_records above
            else:
                rec[r0:r0 + k, 2:2 + min(T, ids.shape[1])] = ids[:, :T]
            r0 += k
        dist.all_gather_into_tensor(gathered, rec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    torch.cuda.set_stream(eng.stream)                      # the engine's own (capturable) stream
    for _ in range(max(args.warmup, 3)):
        gather(eng.step_resident(prep, method))
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.mark_begin()
    for a, b in ev:
        flush.zero_()                                       # evict L2 between timed iterations (untimed)
        a.record()
        gather(eng.step_resident(prep, method))
        b.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = (eng.launches - launches0) // max(1, args.steps)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * BATCH * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end through the public API (e2e) ----------------
    # every step: the step's crops go pinned host -> device, the recognised strings come back to
    # Python.  Two batches are kept in flight with the engine's submit()/collect() pair (the upload
    # and the host-side string decoding of one batch overlap the kernels of the other);
    # `e2e_sync` is the same through the blocking one-call form recognize_packed().
    for _ in range(2):
        eng.recognize_packed(buf, ent, method)
    # a serving process freezes its start-up heap: without this, one generation-2 collection (torch keeps
    # ~10^6 objects alive) lands in some timed iteration and costs 40 ms (seen as e2e between 22 k and 130 k
    # lines/s from run to run; iter_ms_p50_p95_max in the line shows the spread)
    gc.collect()
    gc.freeze()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = eng.recognize_packed(buf, ent, method)
        gather([])                                           # strings are already on the host
    barrier()
    sync_dt = time.perf_counter() - t0
    barrier()
    # warm-up of the two-in-flight form itself: its first iterations grow the caching allocator (two sets of
    # encoder outputs alive at once; a cudaMalloc inside submit() was the 10-90 ms outlier of earlier runs)
    tk = eng.submit(buf, ent, method)
    for _ in range(max(args.warmup, 3)):
        tk2 = eng.submit(buf, ent, method)
        eng.collect(tk)
        tk = tk2
    eng.collect(tk)
    barrier()
    t0 = time.perf_counter()
    iter_s, iter_parts = [], []
    tk = eng.submit(buf, ent, method)
    for _ in range(args.steps - 1):
        ti = time.perf_counter()
        tk2 = eng.submit(buf, ent, method)
        ta = time.perf_counter()
        tk["done"].synchronize()                             # the wait collect() would do, timed separately
        tb = time.perf_counter()
        res = eng.collect(tk)
        gather([])
        tk = tk2
        tc = time.perf_counter()
        iter_s.append(tc - ti)
        iter_parts.append((ta - ti, tb - ta, tc - tb, [b - a for a, b in zip(tk2["marks"][:-1], tk2["marks"][1:])]))
    res = eng.collect(tk)
    gather([])
    barrier()
    e2e_dt = time.perf_counter() - t0
    assert all(r is not None for r in res)
    e2e_t = torch.tensor([e2e_dt, sync_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * args.steps / float(e2e_t[0].item())
    e2e_sync_value = world * BATCH * args.steps / float(e2e_t[1].item())
    h2d = int(buf.numel()) + len(ent) * 32
    d2h = sum(g["n"] * (g["Wb"] // 4 + 2) * 4 for g in prep["groups"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- per-stage profile + roofline of the dominant kernel ----------------
    pk = peaks()
    for _ in range(2):
        eng.step_resident(prep, method)
    torch.cuda.synchronize()
    reps = 5
    prof = eng.profile(lambda: [eng.step_resident(prep, method) for _ in range(reps)])
    widths = {g["Wb"]: g["n"] for g in prep["groups"]}
    if prof.get("ff2", (0, 0))[1] and not prof.get("out_proj", (0, 0))[1]:
        prof["encoder_tail"] = prof.pop("ff2")             # the fused tail is timed under the ff2 stage id
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    stages = {}
    for name, (sms, cnt) in prof.items():
        if cnt == 0:
            continue
        ent_ = {"ms_per_step": sms / reps, "launches_per_step": cnt // reps, "share": sms / total_ms}
        if name in STAGE_FLOPS_PER_LINE:
            fl = sum(STAGE_FLOPS_PER_LINE[name](wb) * n for wb, n in widths.items())
            ent_["tflops"] = fl / (sms / reps / 1e3) / 1e12
            ent_["frac_of_peak"] = ent_["tflops"] / pk["bf16_tflops_sustained"]
        if name == "preprocess":
            by = float(sum(c.size for c in crops) + sum(48 * wb * n for wb, n in widths.items()))   # h*w in, 48*Wb u8 out
            ent_["gbs"] = by / (sms / reps / 1e3) / 1e9
            ent_["frac_of_peak"] = ent_["gbs"] / pk["hbm_gbs"]
        if name in STAGE_BYTES_PER_LINE:
            by = sum(STAGE_BYTES_PER_LINE[name](wb) * n for wb, n in widths.items())
            ent_["gbs"] = by / (sms / reps / 1e3) / 1e9
            ent_["frac_of_peak"] = ent_["gbs"] / pk["hbm_gbs"]
        stages[name] = ent_
    top = max((n for n in stages if n in STAGE_FLOPS_PER_LINE), key=lambda n: stages[n]["ms_per_step"])
    st = stages[top]
    flops_launch = sum(STAGE_FLOPS_PER_LINE[top](wb) * n for wb, n in widths.items()) / max(1, st["launches_per_step"])
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic_v2.json")))
        if top in tr:
            traffic, traffic_note = tr[top]["dram_bytes_per_launch"], tr[top]["note"]
    except Exception:
        pass
    kname = "encoder_block_kernel (out_proj + LN + FFN + LN of one layer)" if top == "encoder_tail" else f"gemm_tc_kernel ({top})"
    roof = {"bound": "tensor", "kernel": kname, "achieved": st["tflops"],
            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": st["tflops"] / pk["bf16_tflops_sustained"],
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside the step)",
            "flops_per_launch": flops_launch, "ms_per_launch": st["ms_per_step"] / max(1, st["launches_per_step"]),
            "share_of_step": st["share"]}
    whole = sum(flops_per_line(wb) * n for wb, n in widths.items()) * world
    tensor_frac = whole * args.steps / (ms_total / 1e3) / 1e12 / pk["bf16_tflops_sustained"] / world

    # ---------------- the other decode method of the metric, device-resident, a few steps ----------------
    other = None
    if world == 1:
        om = "decoder" if method == "ctc" else "ctc"
        for _ in range(2):
            eng.step_resident(prep, om)
        torch.cuda.synchronize()
        oev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in oev:
            flush.zero_()
            a.record()
            eng.step_resident(prep, om)
            b.record()
        torch.cuda.synchronize()
        oms = sum(a.elapsed_time(b) for a, b in oev) / len(oev)
        other = {"decode_method": "accurate" if om == "decoder" else "fast", "value": BATCH / (oms / 1e3),
                 "unit": "lines/s", "ms_per_step": oms, "steps": len(oev)}

    # ---------------- CPU baseline (oracle port of the reference algorithm) ----------------
    cb_v, cb_n, cb_dt = cpu_baseline(cfg, tok, sd, crops, method, budget_s=12.0 if method == "ctc" else 20.0)
    out = {
        "metric": METRIC, "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "lines_per_gpu_per_step": BATCH, "groups": widths,
                   "l2": "256 MiB buffer written between timed iterations", "stem_chunk": args.stem_chunk,
                   "weights": "random-init (seed 0), reference state_dict layout"},
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_dt / args.steps * 1e3, "api": "submit()/collect(), two batches in flight, gc.freeze() after warm-up",
                "sync_value": e2e_sync_value, "sync_api": "recognize_packed(), one blocking call per batch",
                "iter_ms_p50_p95_max": [round(float(np.percentile(np.array(iter_s or [0.0]) * 1e3, q)), 3) for q in (50, 95, 100)],
                "worst_iter_ms_submit_wait_collect": [round(v * 1e3, 3) for v in (iter_parts[int(np.argmax(iter_s))][:3] if iter_s else (0, 0, 0))],
                "worst_iter_submit_phases_ms": [round(v * 1e3, 3) for v in (iter_parts[int(np.argmax(iter_s))][3] if iter_s else [])],
                "submit_phases": "upload enqueue | plan | staging + descriptor copy | preprocess launch | encoder launches | CTC + download enqueue",
                "median_iter_ms_submit_wait_collect": [round(float(np.median([p[k] for p in iter_parts] or [0.0])) * 1e3, 3) for k in range(3)]},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "whole_step_tensor_frac": tensor_frac, "stages": stages, "other_method": other,
        "cpu_baseline": {"value": cb_v, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{cb_n} lines of the same workload in {cb_dt:.1f} s, one line at a time, fp32 oracle"},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="fast", choices=["fast", "accurate"])
    ap.add_argument("--width-mode", default="bucketed", choices=["parity", "bucketed", "masked"])
    ap.add_argument("--stem-chunk", type=int, default=64)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
