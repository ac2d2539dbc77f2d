"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

Imports ``/root/reference/kiri_ocr`` (read-only), writes seeded synthetic checkpoints with
``kiri_ocr_b200.fixtures`` into a temp dir, runs the reference's own ``OCR`` class on CPU fp32 and
records its outputs for a small set of line crops and checkpoint variants.  The GPU box has
no ``/root/reference``; it regenerates the same weights/crops from the seeds and compares with
these vectors.  While generating, every value is also cross-checked against ``oracle/`` so a
drifting restatement fails here first.
"""
import hashlib
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
from oracle import decode as OD, model as OM, preprocess as OP   # noqa: E402

from kiri_ocr import OCR as RefOCR                             # noqa: E402
from kiri_ocr.model import compute_ctc_confidence              # noqa: E402
import torch.nn.functional as F                                # noqa: E402

from tests.golden.cases import VARIANTS, golden_crops, page_case, lines_for   # noqa: E402


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_stepwise_logp(ocr, memp, ids):
    """Reference-style full-prefix decoder runs (model.py:465-485), teacher-forced on ``ids``."""
    m, cfg = ocr.model, ocr.cfg
    rows = []
    seq = [1]
    for t in range(len(ids)):
        inp = torch.tensor([seq])
        tgt = m.dec_pos_enc(m.dec_emb(inp))
        L = len(seq)
        causal = torch.triu(torch.ones((L, L), dtype=torch.bool), diagonal=1)
        out = m.dec_ln(m.dec(tgt=tgt, memory=memp, tgt_mask=causal))
        logp = F.log_softmax(m.dec_head(out)[:, -1, :], dim=-1)
        logp = logp + cfg.LM_FUSION_ALPHA * F.log_softmax(m.lm_head(out)[:, -1, :], dim=-1)
        rows.append(logp[0])
        seq.append(int(ids[t]))
    return torch.stack(rows)


def main():
    torch.set_num_threads(8)
    out = {}
    crops = golden_crops()
    page, boxes = page_case()
    tmp = tempfile.mkdtemp(prefix="kiri_golden_")
    for name, kw in VARIANTS.items():
        sd = FX.make_state_dict(CFG(), 202, **kw)
        path = FX.write_checkpoint(os.path.join(tmp, name), sd)
        ocr_fast = RefOCR(model_path=path, device="cpu", decode_method="fast")
        ocr_acc = RefOCR(model_path=path, device="cpu", decode_method="accurate")
        tok, cfg = ocr_fast.tokenizer, ocr_fast.cfg
        n_lines = lines_for(name)
        for i in range(n_lines):
            roi = crops[i]
            pg = np.pad(roi, 5, mode="edge")             # _preprocess_region pads by 5 px
            box = (5, 5, roi.shape[1], roi.shape[0])
            t = ocr_fast._preprocess_region(pg, box, extra_padding=0)
            with torch.inference_mode():
                plane = ((t[0, 0] * 0.5 + 0.5) * 255.0).round().to(torch.uint8).numpy()
                mem = ocr_fast.model.encode(t)
                logits = ocr_fast.model.ctc_head(mem)
                memp = ocr_fast.model.mem_proj(mem)
                conf, text, length = compute_ctc_confidence(logits, tok)
                best = logits[0].argmax(-1).numpy()
            f_text, f_conf = ocr_fast.recognize_region(t)
            a_text, a_conf = ocr_acc.recognize_region(t)
            assert f_text == text

            # ---- oracle cross-check (fails loudly if the restatement drifts) -------------
            roi_o = OP.crop_region(pg, box, 0)
            plane_o = OP.resize_keep_ratio_pad(roi_o)
            assert np.array_equal(plane, plane_o), (name, i, "plane")
            xo = torch.from_numpy(OP.normalise(plane_o))[None, None]
            assert torch.equal(xo, t), (name, i, "normalise")
            mem_o = OM.encode(sd, xo)
            lg_o = OM.ctc_logits(sd, mem_o)
            e_mem = float((mem_o - mem).abs().max())
            e_lg = float((lg_o - logits).abs().max())
            assert e_mem < 2e-4 and e_lg < 2e-3, (name, i, e_mem, e_lg)
            ot, oc, info = OD.recognize_plane(sd, tok, cfg, plane_o, "ctc")
            assert ot == f_text and abs(oc - f_conf) < 1e-5, (name, i, "fast")
            ot, oc, info = OD.recognize_plane(sd, tok, cfg, plane_o, "decoder")
            assert ot == a_text and abs(oc - a_conf) < 1e-5, (name, i, "accurate", ot, a_text, oc, a_conf)
            ids = info["dec_ids"]
            key = f"{name}/{i}"
            out[f"{key}/plane_sha1"] = np.frombuffer(bytes.fromhex(sha(plane)), np.uint8)
            out[f"{key}/frame_ids"] = best.astype(np.int16)
            out[f"{key}/ctc_conf"] = np.float64(f_conf)
            out[f"{key}/len_est"] = np.int32(length)
            out[f"{key}/fast_text"] = np.array(f_text)
            out[f"{key}/acc_text"] = np.array(a_text)
            out[f"{key}/acc_conf"] = np.float64(a_conf)
            out[f"{key}/dec_ids"] = ids.astype(np.int16)
            if name == "hard" and i < 2:
                out[f"{key}/plane"] = plane
                out[f"{key}/mem"] = mem[0].numpy().astype(np.float32)
                out[f"{key}/ctc_logits"] = logits[0].numpy().astype(np.float32)
            if i < 2:
                with torch.inference_mode():
                    ref_rows = ref_stepwise_logp(ocr_acc, memp, ids)
                _, _, rows = OD.greedy_decode(sd, OM.mem_proj(sd, mem_o), cfg, tok.unk_id + 3, info["len_est"],
                                              forced=list(ids), return_logp=True)
                # oracle rows carry penalties; compare on the unpenalised majority of entries
                d = (rows[: len(ids)] - ref_rows).abs()
                frac_close = float((d < 1e-3).float().mean())
                assert frac_close > 0.97, (name, i, frac_close)
                out[f"{key}/step_logp"] = ref_rows.numpy().astype(np.float32)[:, :]
            print(f"{key}: fast={f_text[:30]!r} conf={f_conf:.4f} len={length} | acc={a_text[:30]!r} "
                  f"conf={a_conf:.4f} steps={len(ids)} | oracle err mem={e_mem:.1e} logits={e_lg:.1e}")
        if name == "hard":
            # page/box path incl. clamping and an empty crop (core.py:506-517)
            for j, box in enumerate(boxes):
                t = ocr_fast._preprocess_region(page, box, extra_padding=5)
                po = OP.preprocess_region(page, box)
                if t is None:
                    assert po is None
                    out[f"page/{j}/plane_sha1"] = np.zeros(20, np.uint8)
                    continue
                plane = ((t[0, 0] * 0.5 + 0.5) * 255.0).round().to(torch.uint8).numpy()
                assert np.array_equal(plane, po), ("page", j)
                out[f"page/{j}/plane_sha1"] = np.frombuffer(bytes.fromhex(sha(plane)), np.uint8)
    dst = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
