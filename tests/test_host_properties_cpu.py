"""Property tests (hypothesis) of the host-side logic around the hot path: width-group planning, the decode slot table,
the sharding cuts, the record format of the exchange and the vectorised text decoding.  CPU only, no compute calls."""
import json
import os
import tempfile

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from kiri_ocr_b200.config import CFG, CharTokenizer

PROFILE = dict(max_examples=60, deadline=None)


def python_target_width(w: int, h: int, H: int = 48) -> int:
    """kiri_ocr/model.py:316-331: nw = max(1, int(round(iw * H / ih))) with Python's round-half-even."""
    return max(1, int(round(w * H / h)))


@settings(**PROFILE)
@given(st.lists(st.tuples(st.integers(1, 3000), st.integers(1, 200)), min_size=1, max_size=300), st.sampled_from(["bucketed", "parity", "masked"]))
def test_plan_groups_places_every_line_once(sizes, mode):
    from kiri_ocr_b200.engine import plan_groups
    cfg = CFG()
    ent = np.array([(0, w, w, h) for w, h in sizes], np.int64)
    groups = plan_groups(ent, cfg, mode)
    seen = np.concatenate([v[0] for v in groups.values()])
    assert sorted(seen.tolist()) == list(range(len(sizes)))                     # every line exactly once
    for Wb, (idx, d, smem, n_strips) in groups.items():
        assert Wb % 128 == 0 and 128 <= Wb <= cfg.IMG_W
        nw = np.array([python_target_width(*sizes[i]) for i in idx])
        assert np.array_equal(d["nw"], nw)                                      # the reference's rounding
        assert (np.minimum(nw, cfg.IMG_W) <= Wb).all()                          # the resized line fits its bucket
        if mode in ("bucketed", "masked"):                                      # ... and no smaller bucket would hold it
            assert (np.minimum(nw, cfg.IMG_W) > Wb - 128).all()
        else:
            assert Wb == cfg.IMG_W
        assert 0 < smem <= 227 * 1024 and n_strips >= 1
        assert (d["strip_w"] >= 1).all() and (d["strip_w"] <= np.maximum(Wb, 1)).all()


@settings(**PROFILE)
@given(st.integers(1, 3000))
def test_decode_slot_table_properties(B):
    from kiri_ocr_b200.engine import decode_slot_table
    t = decode_slot_table(B)
    assert len(t) % 16 == 0
    assert np.array_equal(t[t >= 0], np.arange(B))
    per = [(t[i:i + 16] >= 0).sum() for i in range(0, len(t), 16)]
    assert min(per) >= 1 and max(per) <= 16
    for i, n in zip(range(0, len(t), 16), per):
        assert (t[i:i + n] >= 0).all() and (t[i + n:i + 16] == -1).all()


@settings(**PROFILE)
@given(st.lists(st.floats(0.0, 5000.0, allow_nan=False), min_size=0, max_size=200), st.integers(1, 8))
def test_shard_bounds_are_a_partition(weights, world):
    from kiri_ocr_b200.dist import shard_bounds
    b = shard_bounds(weights, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(weights)
    for (lo, hi), (lo2, _) in zip(b[:-1], b[1:]):
        assert lo <= hi == lo2
    if weights and sum(weights) > 0:
        # no rank carries more than its fair share plus one line
        tot, mx = sum(weights), max(weights)
        for lo, hi in b:
            assert sum(weights[lo:hi]) <= tot / world + mx + 1e-6


@settings(**PROFILE)
@given(st.lists(st.tuples(st.lists(st.integers(2, 400), max_size=40), st.floats(0.0, 1.0, width=32)), min_size=0, max_size=50),
       st.integers(1, 48), st.randoms(use_true_random=False))
def test_records_round_trip_random(lines, lmax, rnd):
    from kiri_ocr_b200.dist import pack_records, unpack_records
    n = len(lines)
    order = list(range(n))
    rnd.shuffle(order)                                                          # records arrive in any order
    ids = [np.array(lines[i][0], np.int32) for i in order]
    conf = [lines[i][1] for i in order]
    rec = pack_records(np.array(order, np.int64), ids, conf, lmax)
    assert rec.shape == (n, 3 + lmax)
    got_ids, got_conf = unpack_records(rec, n + 2)                               # two lines nobody sent
    for i in range(n):
        assert np.array_equal(got_ids[i], np.array(lines[i][0][:lmax], np.int32))
        assert got_conf[i] == np.float32(lines[i][1])
    assert got_ids[n] is None and got_conf[n] is None and got_ids[n + 1] is None


@pytest.fixture(scope="module")
def tok():
    from kiri_ocr_b200 import fixtures as FX
    cfg = CFG()
    d = tempfile.mkdtemp(prefix="kiri_prop_")
    vp = os.path.join(d, "vocab.json")
    with open(vp, "w", encoding="utf-8") as f:
        json.dump(FX.make_vocab(), f, ensure_ascii=False)
    return CharTokenizer(vp, cfg)


@settings(**PROFILE)
@given(st.lists(st.lists(st.integers(-3, 260), max_size=30), min_size=0, max_size=40), st.sampled_from(["ctc", "dec"]))
def test_decode_batch_equals_per_line(tok, lines, space):
    flat = [i for l in lines for i in l]
    lens = [len(l) for l in lines]
    got = tok.decode_batch(np.array(flat, np.int64), np.array(lens, np.int64), space)
    one = tok.decode_collapsed_ctc if space == "ctc" else tok.decode_dec
    assert got == [one(l) for l in lines]


@settings(**PROFILE)
@given(st.lists(st.integers(-2, 230), max_size=80))
def test_decode_ctc_is_collapse_then_lookup(tok, ids):
    """decode_ctc (kiri_ocr/model.py:109-124): equal neighbours count once BEFORE blanks are dropped (so a b a b with a
    blank between two equal ids keeps both), ids < 2, <unk> and out-of-range ids emit nothing."""
    out, prev = [], None
    for i in ids:
        if i == prev:
            continue
        prev = i
        if i < 2 or i >= tok.ctc_classes:
            continue
        t = tok.id_to_token[i - 2]
        if t != tok.unk_token:
            out.append(t)
    assert tok.decode_ctc(ids) == "".join(out)


@settings(**PROFILE)
@given(st.integers(-5, 400))
def test_dec_to_ctc_id_total(tok, dec_id):
    """kiri_ocr/model.py:137-144: specials -> blank, vocabulary ids shift by one, anything else -> <unk>."""
    c = tok.dec_to_ctc_id(dec_id)
    assert 0 <= c < tok.ctc_classes
    if 3 <= dec_id < tok.dec_vocab:
        assert c == dec_id - 1
    elif 0 <= dec_id < 3:
        assert c == 0
    else:
        assert c == tok.unk_id + 2
