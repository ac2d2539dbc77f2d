"""Oracle parity on the EXACT BASELINE workloads (-m gpu): the 256 bucketed line crops bench.py times
(``FX.make_line_crops(256, seed=1234)``, ``width_mode="bucketed"``), with the bench's default-init weights (seed 0)
and with the hardened weights, through the public ``recognize_crops`` for

  * configs[1] "fast"      (CTC greedy,                      kiri_ocr/model.py:343-373, 672-686),
  * configs[2] "accurate"  (greedy KV-cached decoder,        kiri_ocr/model.py:390-600 at BEAM = 1),
  * configs[3] "beam"      (BEAM 5 + CTC forward rescoring,  kiri_ocr/model.py:390-668).

Every line is compared with the oracle run on the plane the reference would see with ``cfg.IMG_W = Wb`` (a width
group of the bucketed mode equals that reference configuration, core.py:430-431).  This covers what the small
fixtures cannot: 16 decode clusters incl. the second-wave cluster, the longest-first ``line_perm``, per-line
``mem_len`` over five width groups in one token stream.  Rule (north_star): ids bit-exact wherever the oracle's
top-1 margin exceeds 2 x the stated tolerance (tests/tolerances.py); on a near tie the device may pick any class
inside the band; everything integer (collapse, lengths, stop rule) is exact given the device's own ids.
"""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from kiri_ocr_b200 import fixtures as FX  # noqa: E402
from kiri_ocr_b200.config import CFG  # noqa: E402
from tests.golden.cases import VARIANTS  # noqa: E402
from tests.test_engine_gpu import _report  # noqa: E402
from tests.tolerances import dec_tol, logit_tol  # noqa: E402

N_LINES = 256


@pytest.fixture(scope="module")
def workload(tok_cfg):
    """Engines + per-line oracle tensors of the bench workload, computed once per weight set."""
    from kiri_ocr_b200.engine import BatchedRecognizer
    from oracle import decode as OD, model as OM, preprocess as OP
    tok, cfg = tok_cfg
    crops = FX.make_line_crops(N_LINES, seed=1234)
    cache = {}

    def get(name):
        if name in cache:
            return cache[name]
        sd = FX.make_state_dict(CFG(), 202, **VARIANTS[name])
        eng = BatchedRecognizer(sd, cfg, tok, width_mode="bucketed")
        _, ent = eng.pack_crops(crops)
        wb = np.zeros(N_LINES, np.int64)
        for Wb, (idx, _, _, _) in eng.plan(ent).items():
            wb[idx] = Wb
        torch.set_num_threads(max(1, torch.get_num_threads()))
        lines = []
        for c, Wb in zip(crops, wb):
            plane = OP.preprocess_crop(c, 48, int(Wb))
            mem = OM.encode(sd, torch.from_numpy(OP.normalise(plane))[None, None])
            lg = OM.ctc_logits(sd, mem)[0]
            best, collapsed, conf, length = OD.ctc_greedy(lg.numpy())
            lines.append({"mem": mem, "logits": lg, "best": best, "collapsed": collapsed, "conf": conf, "length": length})
        cache[name] = (eng, sd, wb, lines)
        return cache[name]
    return crops, get


def _collapse(frames):
    return [int(a) for k, a in enumerate(frames) if a >= 2 and (k == 0 or a != frames[k - 1])]


@pytest.mark.parametrize("name", ["default", "hard"])
def test_config2_fast_256_bucketed(workload, tok_cfg, name):
    tok, cfg = tok_cfg
    crops, get = workload
    eng, sd, wb, lines = get(name)
    tol = logit_tol(sd)
    res = eng.recognize_crops(crops, "ctc", streaming=True)
    st = {"lines": N_LINES, "groups": {int(w): int((wb == w).sum()) for w in np.unique(wb)}, "logit_tol": tol, "frames": 0,
          "frames_equal": 0, "safe_frames": 0, "safe_equal": 0, "all_safe_lines": 0, "all_safe_lines_equal": 0, "text_equal": 0,
          "max_conf_diff": 0.0}
    for i, (r, o) in enumerate(zip(res, lines)):
        lg = o["logits"].numpy()
        got = np.asarray(r.frame_ids, np.int64)
        assert len(got) == wb[i] // 4, i
        top = lg.max(1)
        margin = top - np.sort(lg, axis=1)[:, -2]
        safe = margin > 2 * tol
        assert np.array_equal(got[safe], o["best"][safe]), i                      # bit-exact on margin-safe frames
        assert np.all(top - lg[np.arange(len(got)), got] <= 2 * tol), i           # near ties stay inside the band
        want_ids = _collapse(got.tolist())                                        # integer work on the device's frames
        assert r.ids.tolist() == want_ids and r.len_est == len(want_ids), i
        assert r.text == tok.decode_collapsed_ctc(want_ids), i
        st["frames"] += len(got); st["frames_equal"] += int((got == o["best"]).sum())
        st["safe_frames"] += int(safe.sum()); st["safe_equal"] += int((got[safe] == o["best"][safe]).sum())
        st["max_conf_diff"] = max(st["max_conf_diff"], abs(r.confidence - o["conf"]))
        text_eq = r.text == tok.decode_ctc(o["best"].tolist())
        st["text_equal"] += int(text_eq)
        if safe.all():
            st["all_safe_lines"] += 1
            st["all_safe_lines_equal"] += int(text_eq and r.ids.tolist() == o["collapsed"].tolist())
    _report(f"config2_fast_256_bucketed/{name}", st)
    assert st["all_safe_lines_equal"] == st["all_safe_lines"]
    assert st["max_conf_diff"] < 0.01


@pytest.mark.parametrize("name", ["default", "hard"])
def test_config3_accurate_256_bucketed(workload, tok_cfg, name):
    from oracle import decode as OD, model as OM
    tok, cfg = tok_cfg
    crops, get = workload
    eng, sd, wb, lines = get(name)
    tol = dec_tol(sd)
    res = eng.recognize_crops(crops, "decoder")
    st = {"lines": N_LINES, "dec_tol": tol, "ids_equal": 0, "text_equal": 0, "divergent": 0, "max_conf_diff_on_equal": 0.0,
          "max_step_logp_err": 0.0, "worst_gap_of_device_choice": 0.0, "steps": 0, "len_est_differs": 0}
    unk = tok.unk_id + tok.dec_offset
    for i, (r, o) in enumerate(zip(res, lines)):
        memp = OM.mem_proj(sd, o["mem"])
        # the device bounds its loop with ITS OWN CTC length estimate (exact given its frames, checked above)
        # (near-tie CTC frames may move it: the fast test proves it is exactly the collapse count of the device's frames)
        st["len_est_differs"] += int(r.len_est != o["length"])
        st["max_len_est_diff"] = max(st.get("max_len_est_diff", 0), abs(r.len_est - o["length"]))
        ids = [int(t) for t in r.ids]
        _, lps, rows = OD.greedy_decode(sd, memp, cfg, unk, r.len_est, forced=ids, return_logp=True)
        assert len(lps) == len(ids), (i, len(lps), len(ids))                      # same stop rule / max_steps
        gap = (rows[:len(ids)].max(dim=1).values - torch.tensor(lps)).numpy()     # 0 where the device took the oracle's top-1
        err = np.abs(r.step_logp - np.asarray(lps, np.float32))
        st["max_step_logp_err"] = max(st["max_step_logp_err"], float(err.max()))
        st["worst_gap_of_device_choice"] = max(st["worst_gap_of_device_choice"], float(gap.max()))
        st["steps"] += len(ids)
        assert gap.max() <= 2 * tol, (i, int(gap.argmax()), float(gap.max()))     # every choice is the top-1 or a near tie
        assert err.max() <= tol, (i, float(err.max()))
        if gap.max() == 0.0 and r.len_est == o["length"]:
            # the device followed the oracle's arg-max everywhere: identical ids, so text and confidence must agree
            st["ids_equal"] += 1
            text_ids = ids[:ids.index(2)] if 2 in ids else ids
            assert r.text == tok.decode_dec(text_ids), i
            o_conf = 0.6 * OD.sequence_confidence(lps) + 0.4 * o["conf"]
            st["max_conf_diff_on_equal"] = max(st["max_conf_diff_on_equal"], abs(r.confidence - o_conf))
            st["text_equal"] += 1
        else:
            st["divergent"] += 1
    _report(f"config3_accurate_256_bucketed/{name}", st)
    assert st["max_conf_diff_on_equal"] < 0.02
    assert st["ids_equal"] > 0


@pytest.mark.parametrize("name", ["default", "hard"])
def test_config4_beam5_256_bucketed(workload, tok_cfg, name):
    """BEAM 5 on the whole batch; every 4th line is checked against the oracle's beam search."""
    from oracle import decode as OD, model as OM
    tok, cfg = tok_cfg
    crops, get = workload
    eng, sd, wb, lines = get(name)
    tol = dec_tol(sd)
    bcfg = copy.copy(cfg)
    bcfg.BEAM = 5
    old = eng.cfg.BEAM
    eng.cfg.BEAM = 5
    try:
        res = eng.recognize_crops(crops, "beam")
    finally:
        eng.cfg.BEAM = old
    assert len(res) == N_LINES and all(r is not None for r in res)
    unk = tok.unk_id + tok.dec_offset
    st = {"lines_checked": 0, "hypothesis_equal": 0, "max_conf_diff_on_equal": 0.0, "max_step_logp_err": 0.0, "gaps": []}
    for i in range(0, N_LINES, 4):
        r, o = res[i], lines[i]
        if r.len_est != o["length"]:
            continue                                   # a near-tie CTC frame changed max_steps: not comparable line by line
        st["lines_checked"] += 1
        memp = OM.mem_proj(sd, o["mem"])
        text, conf, info = OD.beam_decode(sd, memp, o["logits"], tok, bcfg)
        best_score, best_seq = info["scored"][0]
        ids = [int(t) for t in r.ids]
        # the scores the device searched with: teacher-forced oracle log-probs of its winning hypothesis
        _, lps = OD.greedy_decode(sd, memp, cfg, unk, r.len_est, forced=ids)
        assert len(lps) == len(ids), i
        err = float(np.abs(r.step_logp - np.asarray(lps, np.float32)).max()) if ids else 0.0
        st["max_step_logp_err"] = max(st["max_step_logp_err"], err)
        assert err <= tol, (i, err)
        if ids == list(best_seq[1:]):
            st["hypothesis_equal"] += 1
            assert r.text == text, i
            st["max_conf_diff_on_equal"] = max(st["max_conf_diff_on_equal"], abs(r.confidence - conf))
        else:
            L = max(1, len(ids))
            mine = sum(lps) / (L ** cfg.BEAM_LENP) + cfg.CTC_FUSION_ALPHA * OD.ctc_alignment_score(o["logits"], [1] + ids, tok)
            st["gaps"].append(float(best_score - mine))
    gaps = st["gaps"]
    st["median_gap"] = float(np.median(gaps)) if gaps else 0.0
    st["max_gap"] = float(max(gaps)) if gaps else 0.0
    _report(f"config4_beam5_256_bucketed/{name}", st)
    assert st["lines_checked"] >= 32
    assert st["max_conf_diff_on_equal"] < 0.02
    if len(gaps) >= 3:
        assert st["median_gap"] < 0.30, st


def test_live_stream_delivers_tokens_while_the_kernel_is_running(workload):
    """True low-latency streaming: with the 256-line batch in flight, the first tokens of the first line reach the host
    (mapped pinned memory, per-step system-scope fences) BEFORE the decode kernel has finished; the streamed ids of
    every line equal what the batch call returns with the streaming token rule."""
    crops, get = workload
    eng, sd, wb, lines = get("hard")
    batch = eng.recognize_crops(crops, "decoder", streaming=True)
    ent = np.zeros((len(crops), 4), np.int64)
    off = 0
    for i, c in enumerate(crops):
        ent[i] = (off, c.shape[1], c.shape[1], c.shape[0])
        off += c.size
    with torch.cuda.stream(eng.stream):
        tk = eng.submit([np.ascontiguousarray(c) for c in crops], ent, "decoder", streaming=True, live=True)
    gen = eng.live_greedy(tk, 0)
    first = [next(gen) for _ in range(3)]
    still_running = not tk["done"].query()
    got0 = [t[0] for t in first] + [t[0] for t in gen]
    assert still_running, "the whole decode had finished before the third token was read"
    assert got0 == batch[0].ids.tolist()
    for li in (1, 17, 100, 255):
        assert [t[0] for t in eng.live_greedy(tk, li)] == batch[li].ids.tolist(), li
    eng.live_finish(tk)
