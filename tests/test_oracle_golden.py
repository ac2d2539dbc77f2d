"""Pin the CPU oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from kiri_ocr_b200 import fixtures as FX
from kiri_ocr_b200.config import CFG
from oracle import decode as OD, model as OM, preprocess as OP
from tests.golden.cases import VARIANTS, golden_crops, lines_for, page_case


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.fixture(scope="module")
def crops():
    return golden_crops()


def test_preprocess_planes_match_reference(golden, crops):
    for i, roi in enumerate(crops):
        page = np.pad(roi, 5, mode="edge")
        plane = OP.preprocess_region(page, (5, 5, roi.shape[1], roi.shape[0]), extra_padding=0)
        assert np.array_equal(_sha(plane), golden[f"hard/{i}/plane_sha1"]), i
        if i < 2:
            assert np.array_equal(plane, golden[f"hard/{i}/plane"])


def test_page_boxes_clamp_and_empty(golden):
    page, boxes = page_case()
    n_empty = 0
    for j, box in enumerate(boxes):
        plane = OP.preprocess_region(page, box)
        want = golden[f"page/{j}/plane_sha1"]
        if plane is None:
            assert not want.any()
            n_empty += 1
        else:
            assert np.array_equal(_sha(plane), want), j
    assert n_empty == 1          # the box fully outside the page (core.py:516-517)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_fast_and_accurate_match_reference(golden, crops, tok_cfg, name):
    tok, cfg = tok_cfg
    sd = FX.make_state_dict(CFG(), 202, **VARIANTS[name])
    n = lines_for(name) if name in ("hard", "eos") else 2      # keep the CPU suite short
    if name == "hard":
        n = 6
    for i in range(n):
        plane = OP.resize_keep_ratio_pad(OP.crop_region(crops[i], (0, 0, crops[i].shape[1], crops[i].shape[0]), 0))
        key = f"{name}/{i}"
        text, conf, info = OD.recognize_plane(sd, tok, cfg, plane, "ctc")
        assert text == str(golden[f"{key}/fast_text"])
        assert abs(conf - float(golden[f"{key}/ctc_conf"])) < 1e-5
        assert info["len_est"] == int(golden[f"{key}/len_est"])
        text, conf, info = OD.recognize_plane(sd, tok, cfg, plane, "decoder")
        assert text == str(golden[f"{key}/acc_text"])
        assert abs(conf - float(golden[f"{key}/acc_conf"])) < 1e-5
        assert np.array_equal(info["dec_ids"], golden[f"{key}/dec_ids"].astype(np.int32))


def test_encoder_tensors_match_reference(golden):
    sd = FX.make_state_dict(CFG(), 202, **VARIANTS["hard"])
    for i in range(2):
        x = torch.from_numpy(OP.normalise(golden[f"hard/{i}/plane"]))[None, None]
        mem = OM.encode(sd, x)
        logits = OM.ctc_logits(sd, mem)
        assert float((mem[0] - torch.from_numpy(golden[f"hard/{i}/mem"])).abs().max()) < 1e-4
        assert float((logits[0] - torch.from_numpy(golden[f"hard/{i}/ctc_logits"])).abs().max()) < 1e-3
        best, _, _, _ = OD.ctc_greedy(logits[0].numpy())
        assert np.array_equal(best.astype(np.int16), golden[f"hard/{i}/frame_ids"])


def test_eos_fixture_terminates_at_mixed_steps(golden):
    steps = [len(golden[f"eos/{i}/dec_ids"]) for i in range(6)]
    last = [int(golden[f"eos/{i}/dec_ids"][-1]) for i in range(6)]
    assert len(set(steps)) >= 3 and last.count(2) >= 4      # EOS branch reached (model.py:545)


def test_blank_fixture_takes_memory_length_branch(golden):
    for i in range(4):
        assert int(golden[f"blank/{i}/len_est"]) == 0
        assert len(golden[f"blank/{i}/dec_ids"]) == 170      # min(512, int(160*1)+10) (model.py:421-425)


@pytest.mark.parametrize("name", ["wide", "wide_b"])
def test_wide_margin_fixture_matches_reference(tok_cfg, name):
    """The fitted-head fixtures (tests/golden/wide.py): the oracle reproduces what the unmodified reference produced
    on the stored checkpoint (text for fast / accurate, ids, confidences) and the stored margins are real."""
    import os
    from tests.golden.wide import MARGIN, wide_crops, wide_state_dict
    tok, cfg = tok_cfg
    gw = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_wide_v1.npz"), allow_pickle=False)
    sd = wide_state_dict(gw, name)
    crops, wbs = wide_crops(name)
    for i, (c, Wb) in enumerate(zip(crops, wbs)):
        key = f"{name}/{i}"
        assert Wb == int(gw[f"{key}/Wb"])
        plane = OP.preprocess_crop(c, 48, Wb)
        text, conf, info = OD.recognize_plane(sd, tok, cfg, plane, "ctc")
        assert text == str(gw[f"{key}/text"]) and abs(conf - float(gw[f"{key}/fast_conf"])) < 1e-5
        assert info["ctc_ids"].tolist() == gw[f"{key}/ctc_ids"].astype(np.int64).tolist()
        lg = OM.ctc_logits(sd, OM.encode(sd, torch.from_numpy(OP.normalise(plane))[None, None]))[0].numpy()
        srt = np.sort(lg, axis=1)
        assert float((srt[:, -1] - srt[:, -2]).min()) > 0.9 * MARGIN
        if i == 0:                                              # one decoder run per case keeps the CPU suite short
            text, conf, info = OD.recognize_plane(sd, tok, cfg, plane, "decoder")
            assert text == str(gw[f"{key}/text"]) and abs(conf - float(gw[f"{key}/acc_conf"])) < 1e-5
            assert info["dec_ids"].tolist() == gw[f"{key}/dec_ids"].astype(np.int64).tolist()
