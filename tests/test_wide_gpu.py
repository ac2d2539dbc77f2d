"""Wide-margin fixtures: the device must reproduce the UNMODIFIED reference exactly (-m gpu).

``tests/golden/wide.py`` builds checkpoints whose CTC and decoder heads are fitted so that every frame and every
decode step of the fixture lines clears the bf16 tolerance band many times over (margin 10 in fp32).  On these,
north_star's rule "token ids bit-exact wherever the top-1 margin exceeds the tolerance" applies to EVERY frame
and step, so frame ids, collapsed ids, decoder ids, streaming ids and text must EQUAL the reference goldens
(``golden_wide_v1.npz``, produced by the reference ``OCR`` itself), for fast, accurate and beam, in parity and
bucketed width modes.  The measured logit / log-prob errors and the margins are written to the parity report,
and the test itself checks that the margin really exceeds 2 x the measured error (the claim is not vacuous).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from kiri_ocr_b200.config import CFG  # noqa: E402
from tests.golden.wide import MARGIN, WIDE_CASES, wide_crops, wide_state_dict  # noqa: E402
from tests.test_engine_gpu import _report  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("wide", "parity"), ("wide", "bucketed"), ("wide_b", "bucketed")]


@pytest.fixture(scope="module")
def gw():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_wide_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def wide_engines(tok_cfg, gw):
    from kiri_ocr_b200.engine import BatchedRecognizer
    tok, _ = tok_cfg
    cache = {}

    def get(name, mode):
        if (name, mode) not in cache:
            sd = wide_state_dict(gw, name)
            cache[(name, mode)] = (BatchedRecognizer(sd, CFG(), tok, width_mode=mode), sd)
        return cache[(name, mode)]
    return get


@pytest.mark.parametrize("name,mode", CASES)
def test_wide_fast_equals_reference(wide_engines, gw, name, mode):
    """decode_method="fast": every frame id, the collapsed ids, the text and the confidence."""
    from oracle import model as OM, preprocess as OP
    eng, sd = wide_engines(name, mode)
    crops, wbs = wide_crops(name)
    res = eng.recognize_crops(crops, "ctc", streaming=True)
    # logits of the same batch through the stage API, for the measured error beside the margin
    buf, ent = eng.pack_crops(crops)
    worst = 0.0
    for Wb, (idx, descs, smem, n_strips) in eng.plan(ent).items():
        planes, _ = eng.preprocess(buf.cuda(), descs, Wb, smem, n_strips)
        lg = eng.encode(planes)["logits"].cpu().numpy()[:, :, :204]
        for j, li in enumerate(idx):
            plane = OP.preprocess_crop(crops[li], 48, Wb)
            assert np.array_equal(planes[j].cpu().numpy(), plane)
            assert Wb == int(gw[f"{name}/{li}/Wb"])                 # parity mode only for the 640-wide case
            ref = OM.ctc_logits(sd, OM.encode(sd, torch.from_numpy(OP.normalise(plane))[None, None]))[0].numpy()
            worst = max(worst, float(np.abs(lg[j] - ref).max()))
    min_margin = min(float(gw[f"{name}/{i}/ctc_margin"]) for i in range(len(crops)))
    for i, r in enumerate(res):
        key = f"{name}/{i}"
        assert np.array_equal(np.asarray(r.frame_ids, np.int64), gw[f"{key}/frame_ids"].astype(np.int64)), key
        assert r.ids.tolist() == gw[f"{key}/ctc_ids"].astype(np.int64).tolist(), key
        assert r.text == str(gw[f"{key}/text"]), key
        assert abs(r.confidence - float(gw[f"{key}/fast_conf"])) < 1e-3, key
        assert r.len_est == len(gw[f"{key}/ctc_ids"])
    _report(f"wide_fast/{name}/{mode}", {"lines": len(crops), "frames": int(sum(len(r.frame_ids) for r in res)),
                                         "logit_maxabs_err": worst, "min_margin": min_margin, "all_equal": True})
    assert min_margin > 4 * worst, (min_margin, worst)               # margin > 2 * (tolerance = 2 * measured error)


@pytest.mark.parametrize("name,mode", CASES)
def test_wide_accurate_equals_reference(wide_engines, gw, tok_cfg, name, mode):
    """decode_method="accurate": decoder ids incl. the EOS step, text, confidence; streaming rule ids."""
    from oracle import decode as OD, model as OM, preprocess as OP
    tok, cfg = tok_cfg
    eng, sd = wide_engines(name, mode)
    crops, wbs = wide_crops(name)
    res = eng.recognize_crops(crops, "decoder")
    stream = eng.recognize_crops(crops, "decoder", streaming=True)
    worst = 0.0
    for i, (r, s) in enumerate(zip(res, stream)):
        key = f"{name}/{i}"
        want = gw[f"{key}/dec_ids"].astype(np.int64).tolist()
        assert r.ids.tolist() == want, key
        assert want[-1] == 2 and r.text == str(gw[f"{key}/text"]), key
        assert abs(r.confidence - float(gw[f"{key}/acc_conf"])) < 0.01, key
        assert s.ids.tolist() == gw[f"{key}/stream_ids"].astype(np.int64).tolist(), key      # raw arg-max rule
        plane = OP.preprocess_crop(crops[i], 48, wbs[i])
        mem = OM.encode(sd, torch.from_numpy(OP.normalise(plane))[None, None])
        _, lps = OD.greedy_decode(sd, OM.mem_proj(sd, mem), cfg, tok.unk_id + 3, r.len_est, forced=want)
        worst = max(worst, float(np.abs(r.step_logp - np.asarray(lps, np.float32)).max()))
    min_margin = min(float(gw[f"{name}/{i}/dec_margin"]) for i in range(len(crops)))
    _report(f"wide_accurate/{name}/{mode}", {"lines": len(crops), "steps": int(sum(len(r.ids) for r in res)),
                                             "step_logp_maxabs_err": worst, "min_margin": min_margin, "all_equal": True})
    assert min_margin > 4 * worst, (min_margin, worst)


@pytest.mark.parametrize("name,mode", CASES)
def test_wide_beam_equals_reference(wide_engines, gw, name, mode):
    """decode_method="beam" (BEAM 3 like the golden, and 5): the winning hypothesis is the reference's."""
    eng, sd = wide_engines(name, mode)
    crops, _ = wide_crops(name)
    old = eng.cfg.BEAM
    try:
        for beam in (3, 5):
            eng.cfg.BEAM = beam
            res = eng.recognize_crops(crops, "beam")
            for i, r in enumerate(res):
                key = f"{name}/{i}"
                assert r.text == str(gw[f"{key}/text"]), (key, beam)
                assert r.ids.tolist() == gw[f"{key}/dec_ids"].astype(np.int64).tolist(), (key, beam)
                if beam == 3:
                    assert abs(r.confidence - float(gw[f"{key}/beam3_conf"])) < 0.01, key
    finally:
        eng.cfg.BEAM = old
    _report(f"wide_beam/{name}/{mode}", {"lines": len(crops), "beams": [3, 5], "all_equal": True})


def test_wide_through_ocr_class(gw, tmp_path):
    """The public ``OCR`` class on the wide checkpoint: text equality for all three decode methods and for the
    character stream (core.py:530-575, 887-1026)."""
    import cv2
    from kiri_ocr_b200 import OCR, fixtures as FX
    name = "wide"
    sd = wide_state_dict(gw, name)
    path = FX.write_checkpoint(str(tmp_path / "ck"), sd)
    crops, _ = wide_crops(name)
    for method in ("fast", "accurate", "beam"):
        ocr = OCR(model_path=path, device="cuda", decode_method=method)
        for i, c in enumerate(crops):
            img = str(tmp_path / f"line{i}.png")
            cv2.imwrite(img, c)
            text, conf = ocr.recognize_single_line_image(img)
            assert text == str(gw[f"{name}/{i}/text"]), (method, i)
            chunks = list(ocr.recognize_streaming(img))
            assert chunks[-1]["finished"] and chunks[-1]["text"] == text
