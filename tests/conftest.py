import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def tok_cfg(tmp_path_factory):
    import json
    from kiri_ocr_b200 import fixtures as FX
    from kiri_ocr_b200.config import CFG, CharTokenizer
    d = tmp_path_factory.mktemp("vocab")
    vp = os.path.join(d, "vocab.json")
    with open(vp, "w", encoding="utf-8") as f:
        json.dump(FX.make_vocab(), f, ensure_ascii=False)
    cfg = CFG()
    return CharTokenizer(vp, cfg), cfg
