"""bench.py — text lines/sec of the batched line-recognition hot path on N B200s.

    python bench.py --gpus 1 --steps 40 --warmup 3                     # configs[1]: 256 bucketed lines, fast (the default line)
    python bench.py --method accurate | --method beam                   # configs[2] / configs[3] on the same crops
    python bench.py --workload pages --gpus N                           # configs[4]: 250 pages x 40 boxes, sharded (strong scaling)
    python bench.py --impl reference [--method ...]                      # the reference algorithm on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload "lines" (default): a "step" is one pass of the hot path (preprocess -> stem -> encoder -> CTC head -> CTC greedy
[-> greedy / beam decoder]) over one batch of 256 synthetic line crops per GPU.  ``value`` is timed with CUDA events with
the packed source crops already in HBM; ``e2e`` goes through the public submit()/collect() pair from pinned host memory
to Python strings.  Every rank works on its own copy of that batch (weak scaling: identical per-GPU work); the path's exchange (fixed-stride result records,
all-gather over NCCL) runs on a side stream so that it overlaps the next step, still inside the timed region.
Workload "pages": 10 000 lines on 250 synthetic pages, page-major sharded over the ranks (strong scaling), ONE all-gather
at the end, and an ordered-equality check of the N-GPU result against a single-GPU run.  One JSON line is printed by rank 0.
"""
import argparse
import gc
import hashlib
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
METHODS = {"fast": "ctc", "accurate": "decoder", "beam": "beam"}


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
    "dec_crosskv": lambda Wb: 2.0 * (Wb // 4) * 256 * 1536,
}
STAGE_BYTES_PER_LINE = {           # algorithmic bytes for the CUDA-core / HBM-bound stages
    "conv1": lambda Wb: 48 * Wb + 48 * Wb * 48 * 2,
    # the frame decisions are taken in the CTC head's epilogue (no logits in HBM): what is left is the collapse stage,
    # 8 bytes per frame in (arg-max id + probability), ids + length + confidence out
    "ctc_greedy": lambda Wb: (Wb // 4) * 8 + 4 * (Wb // 4) + 12,
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


def make_model(beam: int = 0):
    from kiri_ocr_b200 import fixtures as FX
    from kiri_ocr_b200.config import CFG, CharTokenizer
    cfg = CFG()
    if beam:
        cfg.BEAM = beam
    d = tempfile.mkdtemp(prefix="kiri_bench_")
    vp = os.path.join(d, "vocab.json")
    with open(vp, "w", encoding="utf-8") as f:
        json.dump(FX.make_vocab(), f, ensure_ascii=False)
    tok = CharTokenizer(vp, cfg)
    sd = FX.make_state_dict(cfg, tok.vocab_size, seed=0, hardened=False)     # random-init weights
    return cfg, tok, sd


PAGE_HW = (2339, 1654)


def make_pages(p_lo: int, p_hi: int, lines_per_page: int, pinned: bool = True):
    """Synthetic pages [p_lo, p_hi) of the configs[4] document set (seeded per page) in ONE pinned uint8 tensor, and
    their detector boxes (a synthetic TextDetector stand-in: the boxes the pages were drawn with)."""
    from concurrent.futures import ProcessPoolExecutor
    n = p_hi - p_lo
    t = torch.empty((max(n, 1), PAGE_HW[0], PAGE_HW[1]), dtype=torch.uint8)
    if pinned and torch.cuda.is_available():
        t = t.pin_memory()
    boxes = []
    workers = max(1, min(16, (os.cpu_count() or 4) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    with ProcessPoolExecutor(workers) as ex:
        for k, (page, bx) in enumerate(ex.map(_one_page, [(p, lines_per_page) for p in range(p_lo, p_hi)], chunksize=4)):
            t[k] = torch.from_numpy(page)
            boxes.append(bx)
    return t[:n], boxes


def _one_page(a):
    from kiri_ocr_b200 import fixtures as FX
    return FX.make_page(a[1], seed=10_000 + a[0], page_hw=PAGE_HW)


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


def workload_name(args):
    if args.workload == "pages":
        return (f"full-page pipeline: {args.pages} synthetic pages ({PAGE_HW[1]}x{PAGE_HW[0]} px) x {args.lines_per_page} detector boxes "
                f"(synthetic TextDetector stand-in) = {args.pages * args.lines_per_page} lines, recognition sharded page-major over the "
                f"GPUs, random-init kiri recognizer V=202, decode_method={args.method}, width_mode={args.width_mode}")
    return (f"batch {BATCH} synthetic line crops per GPU (heights 24-96 px, widths bucketed to "
            f"{{128,256,384,512,640}} at H=48), random-init kiri recognizer V=202, decode_method={args.method}, "
            f"width_mode={args.width_mode}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kiri_ocr_b200 import fixtures as FX
    cfg, tok, sd = make_model(5 if args.method == "beam" else 0)
    method = METHODS[args.method]
    from oracle import decode as OD, preprocess as OP
    if args.workload == "pages":
        pg, bx = make_pages(0, 2, args.lines_per_page, pinned=False)
        planes = [OP.preprocess_region(pg[p].numpy(), b) for p in range(2) for b in bx[p]]
    else:
        planes = [OP.preprocess_crop(c) for c in FX.make_line_crops(BATCH, seed=1234)[:64]]
    per_step = {"ctc": 16, "decoder": 4, "beam": 2}[method]
    torch.set_num_threads(os.cpu_count() or 1)
    k = 0

    def step():
        nonlocal k
        for _ in range(per_step):
            OD.recognize_plane(sd, tok, cfg, planes[k % len(planes)], method)
            k += 1
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f"{per_step} lines/step of the same workload, one line at a time (reference mode), fp32, "
              f"oracle port with KV-cached decoder and a single encode per line (resample + recognise)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "lines/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "pages" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "decode_method": args.method, "lines_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class Exchange:
    """The path's exchange step for the weak-scaling workload: fixed-stride records {n, conf bits, ids[T]} of this
    rank's batch to every rank (NCCL all-gather).  It runs on a SIDE stream behind an event fence with two record
    buffers, so step i's gather overlaps step i+1's kernels; `finish()` joins it before the timed region ends."""

    def __init__(self, world, T, device):
        import torch.distributed as dist
        self.dist, self.world, self.T = dist, world, T
        self.side = torch.cuda.Stream(device=device)
        self.rec = [torch.zeros((BATCH, 2 + T), dtype=torch.int32, device=device) for _ in range(2)]
        self.out = [torch.empty((world * BATCH, 2 + T), dtype=torch.int32, device=device) for _ in range(2)] if world > 1 else None
        self.i = 0
        self.work = [None, None]
        self.packed = None

    def exchange(self, fill, inputs=()):
        """Called on the main stream right after the step's last kernel: the side stream waits for it (event fence),
        `fill(rec)` builds the step's records there (so the main stream goes straight on to the next step), then the
        all-gather runs on NCCL's stream.  Buffer b is reused two steps later, after its gather."""
        b = self.i & 1
        self.i += 1
        main = torch.cuda.current_stream()
        if self.packed is not None:
            main.wait_event(self.packed)                         # the previous step's records are built (long done): buffers
        self.side.wait_stream(main)                              # they read (ping-pong result slots) may be overwritten
        with torch.cuda.stream(self.side):
            for t in inputs:
                t.record_stream(self.side)                       # allocated on the main stream, read on the side stream
            if self.work[b] is not None:
                self.work[b].wait()                              # side stream waits for gather i-2 of this buffer
                self.work[b] = None
            fill(self.rec[b])
            self.packed = torch.cuda.Event()
            self.packed.record()
            if self.world > 1:
                self.work[b] = self.dist.all_gather_into_tensor(self.out[b], self.rec[b], async_op=True)

    def finish(self):
        for b in (0, 1):
            if self.work[b] is not None:
                self.work[b].wait()
                self.work[b] = None
        torch.cuda.current_stream().wait_stream(self.side)


def run_lines(args):
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
    cfg, tok, sd = make_model(5 if args.method == "beam" else 0)
    eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode=args.width_mode, stem_chunk=args.stem_chunk)
    method = METHODS[args.method]
    crops = FX.make_line_crops(BATCH, seed=1234)          # the SAME synthetic batch on every rank: identical per-GPU work
    buf, ent = eng.pack_crops(crops)
    prep = eng.prepare_resident(buf, ent)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    T = cfg.IMG_W // 4
    torch.cuda.set_stream(eng.stream)                      # the engine's own (capturable) stream
    xch = Exchange(world, T, eng.device)

    def resident_step():
        """One device-resident pass + the exchange of its records.  "beam" has no host-free form (its final ranking
        runs in Python floats like the reference): it goes through submit()/collect() on the device-resident source."""
        if method == "beam":
            tk = eng.submit(prep["src"], ent, "beam")
            res = eng.collect(tk)
            xch.exchange(lambda rec: eng.ticket_records(tk, T, out=rec), tk["keep"])
            return res
        outs = eng.step_resident(prep, method)
        if method == "ctc":
            ids, n, conf = outs[0]

            def fill(rec):
                _lib.check(eng.lib.kiri_pack_records(ids.data_ptr(), n.data_ptr(), conf.data_ptr(), prep["mem_row0"].data_ptr(),
                                                     BATCH, T, rec.data_ptr(), _lib.stream_ptr()), "kiri_pack_records")
                eng.launches += 1
            xch.exchange(fill, (ids, n, conf))
        else:
            d_ids, n_out, sum_lp, _ = outs[0]

            def fill(rec):
                k = min(T, d_ids.shape[1])
                rec[:, 0] = n_out
                rec[:, 1] = sum_lp.view(torch.int32)
                rec[:, 2:2 + k] = d_ids[:, :k]
                eng.launches += 3
            xch.exchange(fill, (d_ids, n_out, sum_lp))
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    for _ in range(max(args.warmup, 3)):
        resident_step()
    xch.finish()
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    tail = torch.cuda.Event(enable_timing=True)
    # the start-up heap is frozen before ANY timed loop (see the e2e leg): a generation-2 collection landing between a
    # step's two events stalls the launching thread for 40-120 ms and the idle GPU time is counted (r02_final run:
    # accurate 18.4 ms per step instead of 6.0 with every kernel at its usual duration)
    gc.collect()
    gc.freeze()
    barrier()
    sampler.mark_begin()
    for a, b in ev:
        flush.zero_()                                       # evict L2 between timed iterations (untimed)
        a.record()
        resident_step()
        b.record()
    xch.finish()                                            # every gather has completed: the tail is part of the time
    tail.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = (eng.launches - launches0) // max(1, args.steps)
    step_ms = sorted(a.elapsed_time(b) for a, b in ev)
    ms = sum(step_ms) + ev[-1][1].elapsed_time(tail)
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * BATCH * args.steps / (ms_total / 1e3)
    n_steps_dec = None
    if method == "decoder":
        outs = eng.step_resident(prep, method)
        torch.cuda.synchronize()
        n_steps_dec = outs[0][1].cpu().numpy().astype(np.int64)            # decode steps per line (slot order)

    # ---------------- end-to-end through the public API (e2e) ----------------
    # every step: the step's crops go pinned host -> device, the recognised strings come back to Python.  Two batches
    # are kept in flight with the engine's submit()/collect() pair (the upload and the host-side string decoding of one
    # batch overlap the kernels of the other); the exchange gathers the REAL records of every batch (built on the device
    # from the ticket's outputs); `e2e_sync` is the same through the blocking one-call form recognize_packed().
    for _ in range(2):
        eng.recognize_packed(buf, ent, method)
    # a serving process freezes its start-up heap: without this, one generation-2 collection (torch keeps
    # ~10^6 objects alive) lands in some timed iteration and costs 40 ms
    gc.collect()
    gc.freeze()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tk = eng.submit(buf, ent, method)
        if method != "beam":
            xch.exchange(lambda rec, tk=tk: eng.ticket_records(tk, T, out=rec), tk["keep"])
        res = eng.collect(tk)
    xch.finish()
    barrier()
    sync_dt = time.perf_counter() - t0
    barrier()
    tk = eng.submit(buf, ent, method)
    for _ in range(max(args.warmup, 3)):
        tk2 = eng.submit(buf, ent, method)
        eng.collect(tk)
        tk = tk2
    eng.collect(tk)
    barrier()
    t0 = time.perf_counter()
    iter_s, iter_parts = [], []

    def sub():
        t = eng.submit(buf, ent, method)
        if method != "beam":
            xch.exchange(lambda rec, t=t: eng.ticket_records(t, T, out=rec), t["keep"])
        return t
    tk = sub()
    for _ in range(args.steps - 1):
        ti = time.perf_counter()
        tk2 = sub()
        ta = time.perf_counter()
        tk["done"].synchronize()                             # the wait collect() would do, timed separately
        tb = time.perf_counter()
        res = eng.collect(tk)
        tk = tk2
        tc = time.perf_counter()
        iter_s.append(tc - ti)
        iter_parts.append((ta - ti, tb - ta, tc - tb, [b - a for a, b in zip(tk2["marks"][:-1], tk2["marks"][1:])]))
    res = eng.collect(tk)
    xch.finish()
    barrier()
    e2e_dt = time.perf_counter() - t0
    assert all(r is not None for r in res)
    e2e_t = torch.tensor([e2e_dt, sync_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * args.steps / float(e2e_t[0].item())
    e2e_sync_value = world * BATCH * args.steps / float(e2e_t[1].item())
    h2d = int(buf.numel()) + len(ent) * 48
    d2h = int(tk["res_words"]) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- per-stage profile + roofline of the dominant kernel ----------------
    pk = peaks()
    prof_method = "ctc" if method == "beam" else method
    for _ in range(2):
        eng.step_resident(prep, prof_method)
    torch.cuda.synchronize()
    reps = 5
    prof = eng.profile(lambda: [eng.step_resident(prep, prof_method) for _ in range(reps)])
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
    traffic_tab = {}
    for fn in ("r02_ncu_traffic.json", "r01_ncu_traffic_v2.json"):
        try:
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", fn)))
            traffic_tab["_file"] = fn
            break
        except Exception:
            pass
    if method == "decoder" and "dec_step" in stages:
        # the accurate path is dominated by the persistent decode kernel: an HBM / L2 streaming problem.  Algorithmic
        # bytes of one launch = every line re-reads its cross K/V (layers x 2 x T x 256 bf16) once per decode step.
        st = stages["dec_step"]
        Ts = np.concatenate([np.full(g["n"], g["Wb"] // 4) for g in prep["groups"]])
        kv_bytes = float((n_steps_dec * Ts).sum()) * cfg.DEC_LAYERS * 2 * 256 * 2
        w_bytes = float(n_steps_dec.max()) * ((BATCH + 15) // 16) * 6.3e6       # weights streamed from L2: steps x clusters x 6.3 MB
        gbs = kv_bytes / (st["ms_per_step"] / 1e3) / 1e9
        tr = traffic_tab.get("dec_step", {})
        roof = {"bound": "hbm", "kernel": "dec_fused_kernel (whole greedy decode of the batch, persistent clusters)",
                "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                "peak_source": f"{pk['src']} HBM copy bandwidth", "bytes_per_launch": kv_bytes,
                "l2_weight_bytes_per_launch": w_bytes, "l2_weight_gbs": w_bytes / (st["ms_per_step"] / 1e3) / 1e9,
                "ms_per_launch": st["ms_per_step"], "share_of_step": st["share"],
                "decode_steps_max_mean": [int(n_steps_dec.max()), float(n_steps_dec.mean())],
                "steps_per_s": float(n_steps_dec.max()) / (st["ms_per_step"] / 1e3)}
    else:
        top = max((n for n in stages if n in STAGE_FLOPS_PER_LINE), key=lambda n: stages[n]["ms_per_step"])
        st = stages[top]
        flops_launch = sum(STAGE_FLOPS_PER_LINE[top](wb) * n for wb, n in widths.items()) / max(1, st["launches_per_step"])
        tr = traffic_tab.get(top, {})
        kname = "encoder_block_kernel (out_proj + LN + FFN + LN of one layer)" if top == "encoder_tail" else f"gemm_tc_kernel ({top})"
        roof = {"bound": "tensor", "kernel": kname, "achieved": st["tflops"],
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": st["tflops"] / pk["bf16_tflops_sustained"],
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside the step)",
                "flops_per_launch": flops_launch, "ms_per_launch": st["ms_per_step"] / max(1, st["launches_per_step"]),
                "share_of_step": st["share"]}
    whole = sum(flops_per_line(wb) * n for wb, n in widths.items()) * world
    tensor_frac = whole * args.steps / (ms_total / 1e3) / 1e12 / pk["bf16_tflops_sustained"] / world

    # ---------------- the other decode method of the metric, device-resident, a few steps ----------------
    other = None
    if world == 1 and method != "beam":
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
                   "weights": "random-init (seed 0), reference state_dict layout",
                   "exchange": "kiri_pack_records + all_gather_into_tensor of the step's records on a side stream (event fence, "
                               "two buffers); the last gather's tail is inside the timed region",
                   **({"beam": 5, "value_path": "submit()/collect() on the device-resident source (the beam's final ranking is "
                       "host-side Python floats like the reference)"} if method == "beam" else {})},
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_dt / args.steps * 1e3, "api": "submit()/collect(), two batches in flight, gc.freeze() after warm-up",
                "sync_value": e2e_sync_value, "sync_api": "recognize_packed(), one blocking call per batch",
                "iter_ms_p50_p95_max": [round(float(np.percentile(np.array(iter_s or [0.0]) * 1e3, q)), 3) for q in (50, 95, 100)],
                "worst_iter_ms_submit_wait_collect": [round(v * 1e3, 3) for v in (iter_parts[int(np.argmax(iter_s))][:3] if iter_s else (0, 0, 0))],
                "worst_iter_submit_phases_ms": [round(v * 1e3, 3) for v in (iter_parts[int(np.argmax(iter_s))][3] if iter_s else [])],
                "submit_phases": "upload enqueue | plan | staging + descriptor copy | preprocess launch | encoder launches | CTC (+ decode) + download enqueue",
                "median_iter_ms_submit_wait_collect": [round(float(np.median([p[k] for p in iter_parts] or [0.0])) * 1e3, 3) for k in range(3)]},
        "gpu_launches": int(launches), "clocks": clocks,
        "step_ms_min_p50_max": [round(step_ms[0], 4), round(step_ms[len(step_ms) // 2], 4), round(step_ms[-1], 4)],
        "roofline": roof, "whole_step_tensor_frac": tensor_frac, "stages": stages, "other_method": other,
        "cpu_baseline": {"value": cb_v, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{cb_n} lines of the same workload in {cb_dt:.1f} s, one line at a time, fp32 oracle"},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _digest(results) -> str:
    h = hashlib.sha1()
    for page in results:
        for r in page:
            h.update(repr(r).encode("utf-8"))
    return h.hexdigest()


def run_pages(args):
    """configs[4]: the full-page pipeline, strong scaling.  `value`: every rank's pages resident in HBM, crops taken on
    the device from the page, one all-gather of all records at the end of the pass.  `e2e`: pinned host pages through
    kiri_ocr_b200.dist.recognize_pages_sharded (the public sharded call) to Python strings on every rank."""
    import torch.distributed as dist
    from kiri_ocr_b200 import _lib, dist as KD
    from kiri_ocr_b200.engine import BatchedRecognizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    method = METHODS[args.method]
    if method == "beam":
        raise SystemExit("--workload pages supports --method fast | accurate")
    n_pages, lpp = args.pages, args.lines_per_page
    n_total = n_pages * lpp
    lo, hi = KD.shard_bounds([lpp] * n_pages, world)[rank]
    # rank 0 draws the whole document set (it also runs the single-GPU check), the others only their shard; the pages are
    # drawn by forked worker processes BEFORE CUDA / NCCL are initialised in this one, and pinned afterwards
    g_lo, g_hi = (0, n_pages) if rank == 0 else (lo, hi)
    pages, boxes = make_pages(g_lo, g_hi, lpp, pinned=False)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("KIRI_BENCH_NO_SAMPLER"):
        sampler.start()
    pages = pages.pin_memory()
    cfg, tok, sd = make_model()
    eng = BatchedRecognizer(sd, cfg, tok, device="cuda", width_mode=args.width_mode, stem_chunk=args.stem_chunk)
    my_pages, my_boxes = pages[lo - g_lo:hi - g_lo], boxes[lo - g_lo:hi - g_lo]
    all_boxes = [None] * n_pages
    T = cfg.IMG_W // 4
    torch.cuda.set_stream(eng.stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident (value) ----------------
    ppb = max(1, args.batch_lines // lpp)                    # pages per batch
    dev_pages = my_pages.to(eng.device)
    preps, n_local = [], 0
    for p0 in range(0, len(my_boxes), ppb):
        p1 = min(len(my_boxes), p0 + ppb)
        ents = []
        for k in range(p0, p1):
            e, valid = eng.boxes_to_entries(PAGE_HW, my_boxes[k], page_offset=(k - p0) * PAGE_HW[0] * PAGE_HW[1])
            ents.append(e[valid])
        ent = np.concatenate(ents)
        prep = eng.prepare_resident(dev_pages[p0:p1].reshape(-1), ent)
        prep["order"] = np.concatenate([g["idx"] for g in prep["groups"]])
        preps.append((prep, n_local))
        n_local += len(ent)
    counts = torch.tensor([n_local], dtype=torch.int64, device="cuda")
    if world > 1:
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        n_max = max(int(c.item()) for c in allc)
    else:
        n_max = n_local
    rec = torch.zeros((n_max, 2 + T), dtype=torch.int32, device="cuda")
    gathered = torch.empty((world * n_max, 2 + T), dtype=torch.int32, device="cuda") if world > 1 else rec
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def resident_pass():
        for prep, r0 in preps:
            outs = eng.step_resident(prep, method)
            n = prep["n_lines"]
            if method == "ctc":
                ids, nn, conf = outs[0]
                _lib.check(eng.lib.kiri_pack_records(ids.data_ptr(), nn.data_ptr(), conf.data_ptr(), prep["mem_row0"].data_ptr(),
                                                     n, T, rec[r0:r0 + n].data_ptr(), _lib.stream_ptr()), "kiri_pack_records")
                eng.launches += 1
            else:
                d_ids, n_out, sum_lp, _ = outs[0]
                k = min(T, d_ids.shape[1])
                rec[r0:r0 + n, 0] = n_out
                rec[r0:r0 + n, 1] = sum_lp.view(torch.int32)
                rec[r0:r0 + n, 2:2 + k] = d_ids[:, :k]
                eng.launches += 3
        if world > 1:
            dist.all_gather_into_tensor(gathered, rec)         # the path's ONE exchange step

    for _ in range(max(1, min(args.warmup, 3))):
        resident_pass()
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gc.collect()                                            # (as in run_lines: no generation-2 collection inside a timed step)
    gc.freeze()
    barrier()
    sampler.mark_begin()
    for a, b in ev:
        flush.zero_()
        a.record()
        resident_pass()
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
    value = n_total * args.steps / (ms_total / 1e3)
    del dev_pages, preps

    # ---------------- e2e: the public sharded call, pinned host pages -> strings on every rank ----------------
    class Shard:                                            # pages[p] is only touched for this rank's range [lo, hi)
        def __getitem__(self, s):
            return my_pages[s.start - lo:s.stop - lo] if isinstance(s, slice) else my_pages[s - lo]

        def __len__(self):
            return n_pages
    boxes_all = [boxes[p - g_lo] if g_lo <= p < g_hi else [(0, 0, 1, 1)] * lpp for p in range(n_pages)]   # only the counts matter off-shard

    def sharded():
        if world == 1:
            res = eng.recognize_pages(my_pages, my_boxes, method, batch_lines=args.batch_lines)
            return [[None if r is None else (r.text, r.confidence) for r in page] for page in res]
        return KD.recognize_pages_sharded(eng, Shard(), boxes_all, method, batch_lines=args.batch_lines)
    res = sharded()
    gc.collect()
    gc.freeze()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        res = sharded()
    barrier()
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(e2e_t[0].item())
    digest = _digest(res)
    same_on_ranks = True
    if world > 1:
        dg = [None] * world
        dist.all_gather_object(dg, digest)
        same_on_ranks = all(d == digest for d in dg)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- SURVEY section 4 tier (v): the N-GPU result equals the single-GPU result, in order ----------------
    single = eng.recognize_pages(pages, boxes, method, batch_lines=args.batch_lines)
    single = [[None if r is None else (r.text, r.confidence) for r in page] for page in single]
    # (float confidences travel as fp32 record fields in the sharded path; compare at fp32)
    def norm(rs):
        return [[None if r is None else (r[0], float(np.float32(r[1]))) for r in page] for page in rs]
    ordered_equal = norm(single) == norm(res)
    pk = peaks()
    whole = flops_per_line(cfg.IMG_W) * n_total
    tensor_frac = whole * args.steps / (ms_total / 1e3) / 1e12 / pk["bf16_tflops_sustained"] / world
    cb_lines = [pages[0].numpy()[y - 5:y + h + 5, x - 5:x + w + 5] for (x, y, w, h) in boxes[0][:16]]
    cb_v, cb_n, cb_dt = cpu_baseline(cfg, tok, sd, cb_lines, method, budget_s=10.0)
    out = {
        "metric": METRIC, "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 3)),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "lines_total": n_total, "pages": n_pages, "lines_per_page": lpp,
                   "batch_lines": args.batch_lines, "pages_of_this_rank": [lo, hi],
                   "l2": "256 MiB buffer written between timed passes (a pass also streams ~1 GB of pages / N)",
                   "weights": "random-init (seed 0), reference state_dict layout",
                   "exchange": "ONE all_gather_into_tensor of every rank's records at the end of the pass, inside the timed region"},
        "e2e": {"value": e2e_value, "unit": "lines/s", "h2d_bytes_per_step": int((hi - lo) * PAGE_HW[0] * PAGE_HW[1]),
                "d2h_bytes_per_step": int(n_local * (T + 2) * 4), "ms_per_step": e2e_dt / e2e_steps * 1e3, "steps": e2e_steps,
                "api": "kiri_ocr_b200.dist.recognize_pages_sharded (engine.recognize_pages per rank: whole pinned pages uploaded "
                       "in place, two batches in flight; one all-gather of records; strings on every rank)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "ordered_equal_to_single_gpu": bool(ordered_equal), "identical_on_all_ranks": bool(same_on_ranks), "result_sha1": digest,
        "roofline": {"bound": "tensor", "kernel": "whole pass (see the lines workload for the per-kernel roofline)",
                     "achieved": whole * args.steps / (ms_total / 1e3) / 1e12 / world, "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": tensor_frac, "traffic": None},
        "whole_step_tensor_frac": tensor_frac,
        "cpu_baseline": {"value": cb_v, "unit": "lines/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{cb_n} lines of page 0 in {cb_dt:.1f} s, one line at a time, fp32 oracle"},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="fast", choices=["fast", "accurate", "beam"])
    ap.add_argument("--width-mode", default="bucketed", choices=["parity", "bucketed", "masked"])
    ap.add_argument("--stem-chunk", type=int, default=64)
    ap.add_argument("--workload", default="lines", choices=["lines", "pages"])
    ap.add_argument("--pages", type=int, default=250)
    ap.add_argument("--lines-per-page", type=int, default=40)
    ap.add_argument("--batch-lines", type=int, default=320)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.workload == "pages" else (40 if args.method == "fast" else 10)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "pages":
        run_pages(args)
    else:
        run_lines(args)


if __name__ == "__main__":
    main()
