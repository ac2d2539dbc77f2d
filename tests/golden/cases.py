"""Golden-vector cases shared by ``make_golden.py`` (build container, imports the reference)
and the tests (anywhere, no reference needed): checkpoint variants and input crops, all seeded."""
import numpy as np

from kiri_ocr_b200 import fixtures as FX

VARIANTS = {
    # name: make_state_dict kwargs
    "hard": dict(seed=1, hardened=True),
    "default": dict(seed=0, hardened=False),
    "eos": dict(seed=2, hardened=True, eos_bias=10.0),
    "blank": dict(seed=3, hardened=True, blank_bias=30.0),
}


def lines_for(name):
    return {"hard": 11, "eos": 6}.get(name, 4)


def golden_crops():
    """Crops shared by generator and tests: 6 bucketed lines + edge cases (SURVEY §8c iv)."""
    crops = FX.make_line_crops(6, seed=7)
    rng = np.random.default_rng(99)
    crops.append(FX._draw_line(rng, 40, 900, inverted=False))      # wider than 640 after resize
    crops.append(rng.integers(0, 256, (37, 1), dtype=np.uint8))    # 1-px-wide crop
    crops.append(FX._draw_line(rng, 64, 853, inverted=True))       # dark background
    crops.append(FX._draw_line(rng, 48, 300, inverted=False))      # already H=48 (no vertical pass)
    crops.append(FX._draw_line(rng, 9, 50, inverted=False))        # strong up-scaling
    return crops


def page_case():
    page, boxes = FX.make_page(6, seed=5, page_hw=(400, 700))
    boxes = list(boxes) + [(0, 0, 120, 30), (650, 380, 80, 40), (690, 100, 30, 20), (710, 10, 5, 5)]
    return page, boxes


