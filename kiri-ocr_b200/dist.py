"""Multi-GPU sharding of line crops.

Every line is independent (eval-mode BN, no cross-line state), so the path shards with no
data-path collective: each rank recognises a contiguous, width-balanced range of the global line
index with replicated weights.  The single exchange step is an all-gather of fixed-stride int32
records ``{line_idx, n_ids, confidence bits, ids[Lmax]}`` over ``torch.distributed`` (NCCL on the
B200 box, gloo in the CPU tests); ids are mapped to strings after the gather and the global order
is restored from ``line_idx``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_bounds(weights: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges [lo, hi) of the line index with near-equal total weight per rank
    (weight = resized line width, i.e. work)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def pack_records(line_idx: np.ndarray, ids: List[np.ndarray], conf: Sequence[float], lmax: int) -> torch.Tensor:
    n = len(line_idx)
    rec = np.zeros((n, 3 + lmax), np.int32)
    rec[:, 0] = line_idx
    lens = np.fromiter((min(len(r), lmax) for r in ids), np.int64, n)
    rec[:, 1] = lens
    if n and lens.sum():
        rec[:, 3:][np.arange(lmax)[None, :] < lens[:, None]] = np.concatenate([np.asarray(r[:lmax], np.int32) for r in ids])
    rec[:, 2] = np.asarray(conf, np.float32).view(np.int32)
    return torch.from_numpy(rec)


def all_gather_records(rec: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks' records, padded per rank to the largest shard (padding rows carry line_idx -1)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(int(c.item()) for c in counts))
    padded = torch.full((n_max, rec.shape[1]), -1, dtype=torch.int32, device=rec.device)
    padded[: rec.shape[0]] = rec
    out = torch.empty((world * n_max, rec.shape[1]), dtype=torch.int32, device=rec.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[out[:, 0] >= 0]


def unpack_records(rec: torch.Tensor, n_total: int):
    """-> (ids per line, confidence per line) in global line order; missing lines give None, lines that failed
    on their rank (n_ids == -1) give ids None and confidence NaN."""
    r = rec.cpu().numpy()
    ids: List[Optional[np.ndarray]] = [None] * n_total
    conf: List[Optional[float]] = [None] * n_total
    cf = r[:, 2].copy().view(np.float32)
    for row, li, k, c in zip(r, r[:, 0].tolist(), r[:, 1].tolist(), cf.tolist()):
        if k < 0:
            conf[li] = float("nan")
            continue
        ids[li] = row[3:3 + k]
        conf[li] = c
    return ids, conf


def _record_width(engine, method: str) -> int:
    """ids per record: T_max for CTC, the decode loop's static capacity for the decoders."""
    T = engine.cfg.IMG_W // 4
    if method == "ctc":
        return T
    cap = getattr(engine, "static_step_cap", None)
    return cap(T) if cap else engine.cfg.MAX_DEC_LEN


def _texts(engine, ids, conf, method):
    """(text, confidence) per line from the gathered ids; the texts of all lines in one vectorised pass when the
    tokenizer offers ``decode_batch``."""
    tok = engine.tok
    rows = [r for r in ids if r is not None]
    if method != "ctc":
        rows = [r[: int(np.argmax(r == tok.dec_eos))] if (r == tok.dec_eos).any() else r for r in rows]
    if hasattr(tok, "decode_batch"):
        flat = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        texts = iter(tok.decode_batch(flat, [len(r) for r in rows], "ctc" if method == "ctc" else "dec"))
    elif method == "ctc":
        texts = iter(tok.decode_collapsed_ctc(r.tolist()) for r in rows)
    else:
        texts = iter(tok.decode_dec(r.tolist()) for r in rows)
    out = []
    for row, c in zip(ids, conf):
        if row is None:
            out.append(None if c is None else LineFailed())
        else:
            out.append((next(texts), c))
    return out


def records_to_results(engine, rec: torch.Tensor, n_total: int, method: str):
    """Gathered records -> per line ``(text, confidence)`` / ``None`` (no record: empty crop) / ``LineFailed()``, in global
    line order.  One pass over the record MATRIX (no per-line arrays): the ids of all lines go through the tokenizer's
    ``decode_batch`` in one call; decoder ids are cut before the first EOS.  Same result as ``unpack_records`` + ``_texts``
    (tests/test_dist_cpu.py), ~10x less host time at 10 000 lines."""
    tok = engine.tok
    if not hasattr(tok, "decode_batch"):
        ids, conf = unpack_records(rec, n_total)
        return _texts(engine, ids, conf, method)
    r = rec.cpu().numpy()
    li, k = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)
    cf = r[:, 2].copy().view(np.float32)
    body = r[:, 3:]
    ok = k >= 0
    n = np.where(ok, np.minimum(k, body.shape[1]), 0)
    if method != "ctc":                                  # cut before the first EOS inside the valid prefix
        is_eos = (body == tok.dec_eos) & (np.arange(body.shape[1])[None, :] < n[:, None])
        first = np.where(is_eos.any(axis=1), is_eos.argmax(axis=1), n)
        n = np.minimum(n, first)
    # texts are decoded in RECORD order (no reordering of the id matrix); the line index only places the results
    flat = body[np.arange(body.shape[1])[None, :] < n[:, None]]
    texts = tok.decode_batch(flat, n[ok], "ctc" if method == "ctc" else "dec")              # failed rows hold no ids
    out: List[object] = [None] * n_total
    if ok.all():
        for j, pair in zip(li.tolist(), zip(texts, cf.tolist())):
            out[j] = pair
        return out
    t = iter(texts)
    for j, good, c in zip(li.tolist(), ok.tolist(), cf.tolist()):
        out[j] = (next(t), c) if good else LineFailed()
    return out


class LineFailed:
    """Placeholder of a region that failed on the rank that owned it (see engine.LineError)."""

    def __repr__(self):
        return "LineFailed()"

    def __eq__(self, other):
        return isinstance(other, LineFailed)


def recognize_pages_sharded(engine, pages, boxes_list, method: str = "ctc", group=None, batch_lines: int = 384,
                            texts_on: Optional[int] = None):
    """configs[4]: detector boxes of many pages, recognition sharded PAGE-MAJOR over the ranks (contiguous page
    ranges balanced by line count, so every rank uploads only its own pages), the engine's pipelined multi-page
    path on every rank, then the path's ONE exchange step: an all-gather of the fixed-stride records.  Every rank
    returns, per page and box, ``(text, confidence)`` / ``None`` (empty crop) / ``LineFailed()``, identical on all
    ranks and identical to a single-rank run.  ``pages[p]`` is only touched on the rank that owns page p (the others
    may pass ``None`` there); ``pages`` may also be a pinned uint8 tensor [n, H, W] (zero-copy uploads).
    ``texts_on``: build the Python strings on that rank only (the others take part in the exchange and return None) -
    turning 10 000 records into strings costs a few milliseconds of host time that N - 1 ranks may not need."""
    import torch.distributed as dist
    from .engine import LineError
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_lines = np.array([len(b) for b in boxes_list], np.int64)
    lo, hi = shard_bounds(n_lines, world)[rank]
    local = engine.recognize_pages(pages[lo:hi], boxes_list[lo:hi], method, batch_lines=batch_lines)
    start = np.concatenate([[0], np.cumsum(n_lines)])
    idx, ids, conf = [], [], []
    for p, page_res in zip(range(lo, hi), local):
        for i, r in enumerate(page_res):
            if r is None:
                continue
            idx.append(int(start[p]) + i)
            if isinstance(r, LineError):
                ids.append(None); conf.append(0.0)
            else:
                ids.append(r.ids); conf.append(r.confidence)
    lmax = _record_width(engine, method)
    rec = pack_records(np.asarray(idx, np.int64), [np.zeros(0, np.int32) if r is None else r for r in ids], conf, lmax)
    for k, r in enumerate(ids):
        if r is None:
            rec[k, 1] = -1
    dev = getattr(engine, "device", torch.device("cpu"))
    allrec = all_gather_records(rec.to(dev) if dist.get_backend(group) == "nccl" else rec, group)
    if texts_on is not None and rank != texts_on:
        return None
    flat = records_to_results(engine, allrec, int(start[-1]), method)
    return [flat[int(start[p]):int(start[p + 1])] for p in range(len(boxes_list))]


def recognize_sharded(engine, src: torch.Tensor, entries: np.ndarray, method: str = "ctc", group=None):
    """Recognise ``entries`` across all ranks of the process group; every rank returns the full,
    ordered list of ``(text, confidence)``.  ``engine.recognize_packed`` does the local work."""
    import torch.distributed as dist
    from .engine import target_widths
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nw = np.minimum(target_widths(entries[:, 2], entries[:, 3], engine.cfg.IMG_H), engine.cfg.IMG_W)
    lo, hi = shard_bounds(nw, world)[rank]
    local = engine.recognize_packed(src, entries[lo:hi], method)
    lmax = _record_width(engine, method)
    rec = pack_records(np.arange(lo, hi), [r.ids for r in local], [r.confidence for r in local], lmax)
    dev = getattr(engine, "device", torch.device("cpu"))
    allrec = all_gather_records(rec.to(dev) if dist.get_backend(group) == "nccl" else rec, group)
    return records_to_results(engine, allrec, len(entries), method)
