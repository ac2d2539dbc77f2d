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
