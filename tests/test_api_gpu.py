"""The OCR class API on the B200 engine (-m gpu): constructor/aliases/errors, result and chunk
schemas (kiri_ocr/core.py:778-784, 963-1000), checkpoint files, duck-typed detector."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from kiri_ocr_b200 import fixtures as FX  # noqa: E402
from kiri_ocr_b200.config import CFG  # noqa: E402


class Box:
    def __init__(self, b, conf=0.9):
        self.bbox, self.confidence = tuple(b), conf


class FakeDetector:
    def __init__(self, boxes):
        self.boxes = boxes

    def detect_lines_objects(self, path):
        return [Box(b) for b in self.boxes]

    def detect_words(self, path):
        return [tuple(b) for b in self.boxes]


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    import cv2
    d = str(tmp_path_factory.mktemp("ckpt"))
    sd = FX.make_state_dict(CFG(), 202, seed=1, hardened=True)
    path = FX.write_checkpoint(d, sd)
    page, boxes = FX.make_page(8, seed=3, page_hw=(600, 900))
    img = os.path.join(d, "page.png")
    cv2.imwrite(img, page)
    return path, img, page, boxes, sd


def test_constructor_aliases_and_errors(setup):
    from kiri_ocr_b200 import OCR
    from kiri_ocr_b200._lib import KiriError
    path = setup[0]
    assert OCR(model_path=path, decode_method="fast").decode_method == "ctc"
    assert OCR(model_path=path, decode_method="accurate").decode_method == "decoder"
    with pytest.warns(DeprecationWarning):
        assert OCR(model_path=path, use_beam_search=False).decode_method == "ctc"
    with pytest.raises(ValueError):
        OCR(model_path=path, decode_method="nope")
    with pytest.raises(KiriError):
        OCR(model_path=path, device="cpu")


def test_extract_text_schema_and_oracle_agreement(setup, tok_cfg):
    from kiri_ocr_b200 import OCR
    from oracle import decode as OD, preprocess as OP
    path, img, page, boxes, sd = setup
    tok, cfg = tok_cfg
    ocr = OCR(model_path=path, decode_method="fast")
    ocr._detector = FakeDetector(boxes)
    text, results = ocr.extract_text(img)
    assert len(results) == len(boxes)
    for i, r in enumerate(results, 1):
        assert set(r) == {"box", "text", "confidence", "det_confidence", "line_number"}
        assert r["line_number"] == i and r["box"] == [int(v) for v in boxes[i - 1]]
        plane = OP.preprocess_region(page, boxes[i - 1])
        t, c, _ = OD.recognize_plane(sd, tok, cfg, plane, "ctc")
        assert abs(c - r["confidence"]) < 0.05
    assert text.count("\n") == len([r for r in results]) - 1 or text
    with pytest.raises(ValueError):
        ocr.extract_text("/nonexistent/file.png")
    # single-line helpers
    t1, c1 = ocr.recognize_single_line_image(img)
    assert isinstance(t1, str) and 0.0 <= c1 <= 1.0
    tens = ocr._preprocess_region(page, boxes[0])
    assert tens.shape == (1, 1, 48, 640) and tens.dtype == torch.float32
    assert np.array_equal(ocr._tensor_to_plane(tens).numpy(), OP.preprocess_region(page, boxes[0]))
    t2, c2 = ocr.recognize_region(tens)
    assert t2 == results[0]["text"]


@pytest.mark.parametrize("method", ["fast", "accurate"])
def test_stream_chars_schema(setup, method):
    from kiri_ocr_b200 import OCR
    path, img, page, boxes, sd = setup
    ocr = OCR(model_path=path, decode_method=method)
    ocr._detector = FakeDetector(boxes[:3])
    chunks = list(ocr.extract_text_stream_chars(img))
    starts = [c for c in chunks if c["region_start"]]
    assert len(starts) == 3 and all(c["step"] == 0 and c["token"] == "" for c in starts)
    keys = {"token", "text", "cumulative_text", "region_number", "total_regions", "step", "region_finished",
            "document_finished", "region_start", "box", "det_confidence"}
    for c in chunks:
        assert keys <= set(c)
        if not c["region_start"]:
            assert "confidence" in c
    assert chunks[-1]["document_finished"] is True or method == "accurate"
    # text of each region grows monotonically
    for rn in (1, 2, 3):
        texts = [c["text"] for c in chunks if c["region_number"] == rn and not c["region_start"]]
        assert all(b.startswith(a) for a, b in zip(texts, texts[1:]))
    final = ocr.process_document(img)
    assert [r["box"] for r in final] == [[int(v) for v in b] for b in boxes[:3]]
    stream = list(ocr.extract_text_streaming(img))
    assert len(stream) == 3 and "cumulative_text" in stream[-1] and stream[-1]["total_regions"] == 3


def test_beam_extract_text(setup):
    """decode_method="beam" (BASELINE config 4) through the document API: same result schema."""
    from kiri_ocr_b200 import OCR
    path, img, page, boxes, sd = setup
    ocr = OCR(model_path=path, decode_method="beam")
    ocr.cfg.BEAM = 5
    ocr._detector = FakeDetector(boxes[:2])
    text, results = ocr.extract_text(img)
    assert len(results) == 2 and isinstance(text, str)
    for r in results:
        assert set(r) >= {"box", "text", "confidence", "det_confidence", "line_number"}
        assert 0.0 <= r["confidence"] <= 1.0
    chunks = list(ocr.extract_text_stream_chars(img))
    assert chunks and {c["region_number"] for c in chunks} == {1, 2}


def test_recognize_region_never_inverts(setup, tok_cfg):
    """A caller tensor with a DARK background (e.g. straight from preprocess_pil) must be recognised as it is:
    the reference's recognize_region never inverts (core.py:530-568); only _preprocess_region does."""
    from kiri_ocr_b200 import OCR, _lib
    from oracle import decode as OD, model as OM, preprocess as OP
    from tests.tolerances import logit_tol
    path, img, page, boxes, sd = setup
    tok, cfg = tok_cfg
    ocr = OCR(model_path=path, decode_method="fast")
    light = ocr._preprocess_region(page, boxes[0])
    dark = -light                                                  # 255 - v in plane space
    plane_dark = ocr._tensor_to_plane(dark).numpy()
    assert plane_dark.mean() < 127
    assert np.array_equal(plane_dark, 255 - OP.preprocess_region(page, boxes[0]))
    t_dark, c_dark = ocr.recognize_region(dark)
    o_text, o_conf, _ = OD.recognize_plane(sd, tok, cfg, plane_dark, "ctc")
    assert abs(c_dark - o_conf) < 0.01
    # frame level, under the margin rule
    eng = ocr.model
    ent = np.array([(0, 640, 640, 48, _lib.CROP_NO_INVERT)], np.int64)
    r = eng.recognize_packed([plane_dark.reshape(-1)], ent, "ctc", streaming=True)[0]
    lg = OM.ctc_logits(sd, OM.encode(sd, torch.from_numpy(OP.normalise(plane_dark))[None, None]))[0].numpy()
    srt = np.sort(lg, axis=1)
    safe = (srt[:, -1] - srt[:, -2]) > 2 * logit_tol(sd)
    assert safe.sum() > 40 and np.array_equal(np.asarray(r.frame_ids)[safe], lg.argmax(1)[safe])
    # and WITHOUT the flag the same bytes are treated as a crop and inverted (the _preprocess_region rule)
    r_inv = eng.recognize_packed([plane_dark.reshape(-1)], ent[:, :4], "ctc", streaming=True)[0]
    t_light, c_light = ocr.recognize_region(light)
    assert abs(r_inv.confidence - c_light) < 1e-6 and r_inv.text == t_light


def test_malformed_region_is_isolated(setup):
    """Per-region isolation (core.py:771-791, 873-885, 1011-1026): a region that cannot be processed is dropped by
    process_document, reported with an ``error`` key by the streaming forms; the other regions are unaffected."""
    from kiri_ocr_b200 import OCR
    path, img, page, boxes, sd = setup
    ocr = OCR(model_path=path, decode_method="fast")
    ocr._detector = FakeDetector([boxes[0], ("a", "b", 3, 4), boxes[1], (1, 2, 3)])
    res = ocr.process_document(img)
    assert [r["line_number"] for r in res] == [1, 3]
    ocr._detector = FakeDetector([boxes[0], boxes[1]])
    clean = ocr.process_document(img)
    assert [r["text"] for r in res] == [r["text"] for r in clean]
    ocr._detector = FakeDetector([boxes[0], ("a", "b", 3, 4), boxes[1], (1, 2, 3)])
    st = list(ocr.process_document_streaming(img))
    assert [("error" in r) for r in st] == [False, True, False, True] and st[1]["text"] == "" and st[1]["total_regions"] == 4
    chunks = list(ocr.extract_text_stream_chars(img))
    errs = [c for c in chunks if "error" in c]
    assert [c["region_number"] for c in errs] == [2, 4] and errs[-1]["document_finished"] is True
    assert all(c["region_finished"] and c["region_start"] for c in errs)


def test_recognize_pages_equals_per_page(setup):
    """The pipelined multi-page path (whole pages per batch, two batches in flight) returns exactly what the
    per-page call returns, page by page and box by box, for fast and accurate."""
    from kiri_ocr_b200 import OCR
    path, img, page, boxes, sd = setup
    ocr = OCR(model_path=path, decode_method="fast")
    eng = ocr.model
    pages, bl = [], []
    for s in range(5):
        p, b = FX.make_page(6, seed=20 + s, page_hw=(420, 760))
        pages.append(p)
        bl.append(list(b) + ([(750, 410, 30, 30), (900, 900, 5, 5)] if s == 2 else []))   # a clamped and an empty box
    bl[3] = []                                                     # a page without boxes
    for method in ("ctc", "decoder"):
        multi = eng.recognize_pages(pages, bl, method, batch_lines=12)
        assert [len(m) for m in multi] == [len(b) for b in bl]
        for p, b, m in zip(pages, bl, multi):
            single = eng.recognize_boxes(p, b, method) if b else []
            for x, y in zip(single, m):
                assert (x is None) == (y is None)
                if x is not None:
                    assert x.text == y.text and x.confidence == y.confidence and np.array_equal(x.ids, y.ids)
    assert multi[2][-1] is None


def test_bgr_page_ingest_equals_host_gray(setup, tmp_path):
    """A colour page decoded by cv2.imread goes to the device as BGR and is converted there: results must equal the
    path that converts on the host with cv2 (the reference's own order of operations, core.py:762-766)."""
    import cv2
    from kiri_ocr_b200 import OCR
    path, img, page, boxes, sd = setup
    rng = np.random.default_rng(3)
    colour = np.stack([page, np.clip(page.astype(np.int16) - 9, 0, 255).astype(np.uint8), rng.integers(200, 256, page.shape, dtype=np.uint8)], -1)
    cpath = str(tmp_path / "colour.png")
    cv2.imwrite(cpath, colour)
    ocr = OCR(model_path=path, decode_method="fast")
    ocr._detector = FakeDetector(boxes)
    res = ocr.process_document(cpath)
    bgr = cv2.imread(cpath)
    assert bgr.ndim == 3
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    want = ocr.model.recognize_boxes(gray, boxes, "ctc")
    assert len(res) == len(want)
    for r, w in zip(res, want):
        assert r["text"] == w.text and r["confidence"] == float(w.confidence)
    # mixed batch: a BGR page and a gray page in one multi-page call
    multi = ocr.model.recognize_pages([bgr, gray], [boxes, boxes], "ctc", batch_lines=1000)
    for a, b, w in zip(multi[0], multi[1], want):
        assert a.text == w.text == b.text and a.confidence == w.confidence == b.confidence
