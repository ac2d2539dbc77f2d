"""Greedy attention decoder ("accurate") parity on the B200 (-m gpu).

Free-running greedy decode is chaotic after the first divergent token, so parity is checked
two ways: (1) teacher-forced — the oracle's token sequence is fed and the
per-step penalised log-prob of each fed token must match within dec_tol(sd) (tests/tolerances.py: 0.10 on
the "hard" fixture, 2x the measured 0.042); (2) free-running
— ids / text / confidence against the reference goldens, required identical on the fixtures
whose steps are all margin-safe, reported otherwise (gpurun_out/parity_report.json).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from kiri_ocr_b200 import fixtures as FX  # noqa: E402
from kiri_ocr_b200.config import CFG  # noqa: E402
from tests.golden.cases import VARIANTS, golden_crops, lines_for  # noqa: E402
from tests.test_engine_gpu import _report, engines  # noqa: E402,F401

from tests.tolerances import dec_tol  # noqa: E402


def _encode(eng, crops):
    buf, ent = eng.pack_crops(crops)
    idx, descs, smem, n_strips = eng.plan(ent)[640]
    planes, _ = eng.preprocess(buf.cuda(), descs, 640, smem, n_strips)
    enc = eng.encode(planes)
    ids, n_ids, conf, _, _ = eng.ctc_greedy(enc["logits"])
    return enc, n_ids, conf


@pytest.mark.parametrize("name", ["hard", "eos", "blank"])
def test_teacher_forced_step_logp(engines, golden, tok_cfg, name):
    from oracle import decode as OD, model as OM, preprocess as OP
    tok, cfg = tok_cfg
    eng, sd = engines(name)
    DEC_LOGP_ATOL = dec_tol(sd)
    n = min(4, lines_for(name))
    crops = golden_crops()[:n]
    enc, n_ids, conf = _encode(eng, crops)
    T = 160
    gold = [golden[f"{name}/{i}/dec_ids"].astype(np.int32) for i in range(n)]
    Lmax = max(len(g) for g in gold)
    forced = torch.zeros((n, Lmax), dtype=torch.int32)
    for i, g in enumerate(gold):
        forced[i, :len(g)] = torch.from_numpy(g)
        forced[i, len(g):] = 2
    # the device derives max_steps from ITS OWN length estimate; feed the golden one so both sides stop alike
    len_est = torch.tensor([int(golden[f"{name}/{i}/len_est"]) for i in range(n)], dtype=torch.int32, device="cuda")
    ids, n_out, sum_lp, slp, spr, steps = eng.decode_greedy(enc["mem_bf16"], len_est, n, T, Lmax, forced=forced.cuda(),
                                                            want_steps=True)
    torch.cuda.synchronize()
    worst = 0.0
    for i, c in enumerate(crops):
        plane = OP.preprocess_crop(c)
        x = torch.from_numpy(OP.normalise(plane))[None, None]
        memp = OM.mem_proj(sd, OM.encode(sd, x))
        o_ids, o_lp = OD.greedy_decode(sd, memp, cfg, tok.unk_id + 3, int(len_est[i]), forced=list(gold[i]))
        k = len(gold[i])
        assert int(n_out[i]) == k, (i, int(n_out[i]), k)
        assert np.array_equal(ids[i, :k].cpu().numpy(), gold[i])
        d = np.abs(slp[i, :k].cpu().numpy() - np.asarray(o_lp[:k], np.float32))
        worst = max(worst, float(d.max()))
    _report(f"decoder_forced/{name}", {"max_abs_step_logp_err": worst, "lines": n, "steps": int(steps)})
    assert worst <= DEC_LOGP_ATOL


@pytest.mark.parametrize("name", ["hard", "eos", "blank", "default"])
def test_free_running_accurate_vs_goldens(engines, golden, tok_cfg, name):
    """Free-running greedy decode.  A line whose ids equal the golden must also match its
    confidence.  A line that diverges must diverge at a NEAR TIE: the oracle, teacher-forced on the
    device's own sequence, must rank every device-chosen token within 2*DEC_LOGP_ATOL of its own
    top-1 at that step (so ids are bit-exact wherever the margin exceeds the tolerance)."""
    from oracle import decode as OD, model as OM, preprocess as OP
    tok, cfg = tok_cfg
    eng, sd = engines(name)
    DEC_LOGP_ATOL = dec_tol(sd)
    n = lines_for(name)
    crops = golden_crops()[:n]
    res = eng.recognize_crops(crops, "decoder")
    same_text = same_ids = 0
    first_div = []
    conf_diff = worst_gap = 0.0
    for i, r in enumerate(res):
        g = golden[f"{name}/{i}/dec_ids"].astype(np.int32)
        eq = len(r.ids) == len(g) and np.array_equal(r.ids, g)
        same_ids += int(eq)
        same_text += int(r.text == str(golden[f"{name}/{i}/acc_text"]))
        if eq:
            conf_diff = max(conf_diff, abs(r.confidence - float(golden[f"{name}/{i}/acc_conf"])))
            continue
        m = min(len(r.ids), len(g))
        dv = np.nonzero(r.ids[:m] != g[:m])[0]
        first_div.append(int(dv[0]) if len(dv) else m)
        x = torch.from_numpy(OP.normalise(OP.preprocess_crop(crops[i])))[None, None]
        mem = OM.encode(sd, x)
        _, _, _, length = OD.ctc_greedy(OM.ctc_logits(sd, mem)[0].numpy())
        assert length == int(golden[f"{name}/{i}/len_est"])
        # the device bounds its loop with ITS OWN CTC length estimate, which may differ from the
        # oracle's by the near-tie frames (checked in test_engine_gpu); feed it to the oracle
        assert abs(r.len_est - length) <= 3, (i, r.len_est, length)
        length = r.len_est
        ids = [int(t) for t in r.ids]
        _, lps, rows = OD.greedy_decode(sd, OM.mem_proj(sd, mem), cfg, tok.unk_id + 3, length, forced=ids,
                                        return_logp=True)
        assert len(lps) == len(ids), (i, len(lps), len(ids))      # same stop rule / max_steps
        gap = (rows[:len(ids)].max(dim=1).values - torch.tensor(lps)).numpy()
        worst_gap = max(worst_gap, float(gap.max()))
        assert gap.max() <= 2 * DEC_LOGP_ATOL, (i, int(gap.argmax()), float(gap.max()))
        d = np.abs(r.step_logp - np.asarray(lps, np.float32))
        assert d.max() <= DEC_LOGP_ATOL, (i, float(d.max()))
    _report(f"decoder_free/{name}", {"lines": n, "ids_equal": same_ids, "text_equal": same_text,
                                     "first_divergence_steps": first_div, "max_conf_diff_on_equal": conf_diff,
                                     "worst_oracle_gap_of_device_choice": worst_gap})
    assert conf_diff < 0.02
    if name == "blank":
        assert all(len(r.ids) == 170 for r in res)          # len_ctc == 0 -> 170 steps (model.py:421-425)
    if name == "eos":
        assert sum(int(r.ids[-1]) == 2 for r in res) >= 3  # EOS termination at mixed steps


def test_streaming_rule_raw_argmax(engines, golden, tok_cfg):
    """select_raw=1 picks the arg-max of the RAW dec_head soft-max (model.py:915-917)."""
    from oracle import decode as OD, model as OM, preprocess as OP
    tok, cfg = tok_cfg
    eng, sd = engines("hard")
    crops = golden_crops()[:2]
    res = eng.recognize_crops(crops, "decoder", streaming=True)
    for i, c in enumerate(crops):
        x = torch.from_numpy(OP.normalise(OP.preprocess_crop(c)))[None, None]
        mem = OM.encode(sd, x)
        _, _, _, length = OD.ctc_greedy(OM.ctc_logits(sd, mem)[0].numpy())
        chunks = list(OD.greedy_stream_chunks(sd, OM.mem_proj(sd, mem), tok, cfg, length))
        want = [ch["token_id"] for ch in chunks]
        got = res[i].ids.tolist()
        m = min(len(want), len(got))
        agree = next((k for k, (a, b) in enumerate(zip(want[:m], got[:m])) if a != b), m)      # length of the common prefix
        rec = {"steps": m, "agree_prefix": agree}
        if agree < m:
            # the streamed sequence may only leave the oracle's at a NEAR TIE of the raw dec_head logits: the oracle, fed the
            # common prefix, must rank the device's token within 2 x tolerance of its own choice (bf16 operands on the
            # device; on this random-init fixture the margins are of that size - the wide-margin fixtures of
            # tests/test_wide_gpu.py demand the full sequence)
            st = OM.DecoderState(sd, OM.mem_proj(sd, mem), 8)
            seq = [1] + want[:agree]
            for t_in in seq:
                dec, _ = OM.decoder_step(st, torch.tensor([t_in]))
            gap = float(dec[0, want[agree]] - dec[0, got[agree]])
            rec["first_divergence_gap"] = gap
            _report(f"decoder_stream/hard/{i}", rec)
            assert 0.0 <= gap <= 2 * dec_tol(sd, 0.0), (i, agree, gap)
            continue
        _report(f"decoder_stream/hard/{i}", rec)
        if res[i].len_est == length:                            # same CTC length estimate -> same step bound
            assert len(got) == len(want), (i, len(got), len(want))
