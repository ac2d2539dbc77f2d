"""world_size-2 gloo test of the multi-GPU host logic: contiguous width-balanced sharding, the
single all-gather of fixed-stride records and global order restoration.  The engine is a stub
(deterministic ids from the crop geometry) — the device path is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from kiri_ocr_b200 import dist as KD
from kiri_ocr_b200.config import CFG
from kiri_ocr_b200.engine import LineResult


class StubEngine:
    def __init__(self, tok):
        self.cfg, self.tok, self.device = CFG(), tok, torch.device("cpu")
        self.calls = []

    def recognize_packed(self, src, entries, method="ctc", streaming=False):
        self.calls.append(len(entries))
        out = []
        for off, pitch, w, h in entries:
            ids = np.array([2 + (int(w) % 150), 2 + (int(h) % 150), 2 + (int(off) % 150)][: 1 + int(w) % 3], np.int32)
            out.append(LineResult(self.tok.decode_collapsed_ctc(ids.tolist()), 0.25 + (int(w) % 7) / 10.0, 0.0, ids))
        return out


    def recognize_pages(self, pages, boxes_list, method="ctc", streaming=False, batch_lines=384):
        from kiri_ocr_b200.engine import LineError
        self.calls.append(len(boxes_list))
        out = []
        for page, boxes in zip(pages, boxes_list):
            res = []
            for (x, y, w, h) in boxes:
                if w <= 0:
                    res.append(None)                                              # empty crop: no result
                elif h == 13:
                    res.append(LineError("boom"))                                 # a region that failed on its rank
                else:
                    ids = np.array([2 + (int(x) % 150), 2 + (int(y) % 150), 2 + int(page) % 150][: 1 + int(w) % 3], np.int32)
                    res.append(LineResult(self.tok.decode_collapsed_ctc(ids.tolist()), 0.25 + (int(w) % 7) / 10.0, 0.0, ids))
            out.append(res)
        return out


def _pages_case():
    rng = np.random.default_rng(5)
    n_pages = 9
    boxes = [[(int(rng.integers(0, 500)), int(rng.integers(0, 500)), int(rng.integers(0, 90)), int(rng.integers(10, 40)))
              for _ in range(int(rng.integers(0, 7)))] for _ in range(n_pages)]
    boxes[3][0:0] = [(5, 5, 0, 20), (7, 7, 30, 13)]                               # an empty crop and a failing region
    return list(range(n_pages)), boxes


def _pages_worker(rank, world, port, vocab_path, q, texts_on):
    import torch.distributed as dist
    from kiri_ocr_b200.config import CharTokenizer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = StubEngine(CharTokenizer(vocab_path, CFG()))
    pages, boxes = _pages_case()
    res = KD.recognize_pages_sharded(eng, pages, boxes, "ctc", texts_on=texts_on)
    q.put((rank, eng.calls, res))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, vocab_path, q):
    import torch.distributed as dist
    from kiri_ocr_b200.config import CharTokenizer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tok = CharTokenizer(vocab_path, CFG())
    eng = StubEngine(tok)
    rng = np.random.default_rng(0)
    n = 37
    ent = np.stack([np.arange(n) * 1000, rng.integers(50, 900, n), rng.integers(20, 2000, n), rng.integers(10, 90, n)], 1)
    ent[:, 1] = ent[:, 2]
    res = KD.recognize_sharded(eng, None, ent, "ctc")
    q.put((rank, eng.calls, res))
    dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    w = np.random.default_rng(1).integers(16, 640, 1000)
    for world in (1, 2, 4, 8):
        b = KD.shard_bounds(w, world)
        assert b[0][0] == 0 and b[-1][1] == len(w)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [w[lo:hi].sum() for lo, hi in b]
        assert max(loads) - min(loads) <= 2 * 640
    assert KD.shard_bounds([], 4) == [(0, 0)] * 4
    assert KD.shard_bounds([5.0], 2)[0][1] + (KD.shard_bounds([5.0], 2)[1][1] - KD.shard_bounds([5.0], 2)[1][0]) == 1


def test_records_round_trip():
    ids = [np.array([5, 6, 7], np.int32), np.zeros(0, np.int32), np.arange(2, 162, dtype=np.int32)]
    rec = KD.pack_records(np.array([4, 0, 2]), ids, [0.5, 1.0, 0.123], 160)
    got, conf = KD.unpack_records(rec, 5)
    assert np.array_equal(got[4], ids[0]) and len(got[0]) == 0 and np.array_equal(got[2], ids[2])
    assert got[1] is None and got[3] is None
    assert abs(conf[2] - 0.123) < 1e-7


@pytest.mark.parametrize("method", ["ctc", "decoder"])
def test_records_to_results_equals_the_per_line_form(tok_cfg, method):
    """The vectorised record -> (text, confidence) pass must give exactly what unpack_records + _texts give: lines
    without a record (None), failed lines (n = -1), empty lines, decoder ids cut at the first EOS, unknown ids."""
    tok, _ = tok_cfg
    eng = StubEngine(tok)
    rng = np.random.default_rng(3)
    n, lmax = 500, 64
    lens = rng.integers(0, lmax + 1, n)
    rec = np.zeros((n, 3 + lmax), np.int32)
    rec[:, 0], rec[:, 1] = rng.permutation(n), lens
    for i in range(n):
        rec[i, 3:3 + lens[i]] = rng.integers(0, tok.vocab_size + 6, lens[i])
    rec[::41, 1] = -1
    rec[:, 2] = rng.random(n).astype(np.float32).view(np.int32)
    rec = torch.from_numpy(rec[rng.random(n) > 0.05])
    ids, conf = KD.unpack_records(rec, n)
    want = KD._texts(eng, ids, conf, method)
    got = KD.records_to_results(eng, rec, n, method)
    assert len(got) == n and sum(g is None for g in got) == sum(w is None for w in want) > 0
    for g, w in zip(got, want):
        assert (g is None and w is None) or g == w


def test_two_rank_gloo_gather_restores_order(tok_cfg, tmp_path):
    import json
    from kiri_ocr_b200 import fixtures as FX
    vp = str(tmp_path / "vocab.json")
    json.dump(FX.make_vocab(), open(vp, "w", encoding="utf-8"), ensure_ascii=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, vp, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort()
    (r0, calls0, res0), (r1, calls1, res1) = outs
    assert res0 == res1 and len(res0) == 37 and all(r is not None for r in res0)
    assert calls0[0] + calls1[0] == 37 and min(calls0[0], calls1[0]) >= 10       # both ranks got work
    # identical to the single-process answer, in the original order
    tok, _ = tok_cfg
    eng = StubEngine(tok)
    rng = np.random.default_rng(0)
    n = 37
    ent = np.stack([np.arange(n) * 1000, rng.integers(50, 900, n), rng.integers(20, 2000, n), rng.integers(10, 90, n)], 1)
    ent[:, 1] = ent[:, 2]
    want = [(r.text, r.confidence) for r in eng.recognize_packed(None, ent)]
    for (t, c), (wt, wc) in zip(res0, want):
        assert t == wt and abs(c - wc) < 1e-6


@pytest.mark.parametrize("texts_on", [None, 0])
def test_two_rank_gloo_pages_sharded(tok_cfg, tmp_path, texts_on):
    """recognize_pages_sharded: page-major shards, one exchange, every page's boxes back in order incl. the empty crop
    (None) and the failed region (LineFailed); texts_on=0 builds the strings on rank 0 only."""
    import json
    from kiri_ocr_b200 import fixtures as FX
    vp = str(tmp_path / "vocab.json")
    json.dump(FX.make_vocab(), open(vp, "w", encoding="utf-8"), ensure_ascii=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pages_worker, args=(r, 2, port, vp, q, texts_on)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, calls0, res0), (_, calls1, res1) = outs
    pages, boxes = _pages_case()
    assert calls0[0] + calls1[0] == len(pages) and min(calls0[0], calls1[0]) >= 1
    if texts_on is None:
        assert res0 == res1
    else:
        assert res1 is None
    tok, _ = tok_cfg
    want = StubEngine(tok).recognize_pages(pages, boxes)
    assert len(res0) == len(pages)
    for got_p, want_p in zip(res0, want):
        assert len(got_p) == len(want_p)
        for g, w in zip(got_p, want_p):
            if w is None:
                assert g is None
            elif not isinstance(w, LineResult):
                assert g == KD.LineFailed()
            else:
                assert g[0] == w.text and abs(g[1] - w.confidence) < 1e-6
    assert res0[3][0] is None and res0[3][1] == KD.LineFailed()
