"""Generate tests/golden/golden_beam_v1.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_beam.py

Runs ``kiri_ocr.OCR(decode_method="beam")`` (``/root/reference``, CPU fp32) with ``cfg.BEAM`` in
{3, 5} on the first lines of the seeded golden crops and records text and confidence; every value
is cross-checked against ``oracle.decode.beam_decode`` while generating.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from kiri_ocr_b200 import fixtures as FX                      # noqa: E402
from kiri_ocr_b200.config import CFG                           # noqa: E402
from oracle import decode as OD, preprocess as OP              # noqa: E402
from kiri_ocr import OCR as RefOCR                             # noqa: E402
from tests.golden.cases import VARIANTS, golden_crops          # noqa: E402

BEAM_CASES = {"hard": 4, "eos": 4, "blank": 1}                 # variant -> number of lines


def main():
    torch.set_num_threads(8)
    crops = golden_crops()
    tmp = tempfile.mkdtemp(prefix="kiri_golden_beam_")
    out = {}
    for name, n_lines in BEAM_CASES.items():
        sd = FX.make_state_dict(CFG(), 202, **VARIANTS[name])
        path = FX.write_checkpoint(os.path.join(tmp, name), sd)
        ocr = RefOCR(model_path=path, device="cpu", decode_method="beam")
        for beam in (3, 5):
            ocr.cfg.BEAM = beam
            for i in range(n_lines):
                roi = crops[i]
                pg = np.pad(roi, 5, mode="edge")
                t = ocr._preprocess_region(pg, (5, 5, roi.shape[1], roi.shape[0]), extra_padding=0)
                text, conf = ocr.recognize_region(t)
                cfg = CFG()
                cfg.BEAM = beam
                otext, oconf, info = OD.recognize_plane(sd, ocr.tokenizer, cfg, OP.preprocess_crop(roi), "beam")
                assert otext == text and abs(oconf - conf) < 1e-6, (name, beam, i)
                key = f"{name}/b{beam}/{i}"
                out[f"{key}/text"] = np.array(text)
                out[f"{key}/conf"] = np.float64(conf)
                out[f"{key}/best_ids"] = np.asarray(info["scored"][0][1][1:], np.int32)
                print(key, repr(text[:30]), round(conf, 6))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_beam_v1.npz"), **out)


if __name__ == "__main__":
    main()
