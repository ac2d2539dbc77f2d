"""Generate tests/golden/golden_wide_v1.npz: the wide-margin fixtures (tests/golden/wide.py).

    python tests/golden/make_golden_wide.py            (build container only: imports /root/reference)

For every case: fit ``ctc_head.2`` and ``dec_head`` on the oracle's features, verify the margins with the oracle
(fp32), then run the UNMODIFIED reference ``OCR`` (fast / accurate / beam 3) on the resulting checkpoint (with
``IMG_W`` = the line's batch width in the checkpoint's meta, core.py:430-431) and record its outputs next to the
fitted weights.  The reference must spell exactly the constructed text: the fixture is self-checking.
"""
import copy
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from kiri_ocr_b200 import fixtures as FX                      # noqa: E402
from kiri_ocr_b200.config import CFG                           # noqa: E402
from oracle import decode as OD, model as OM, preprocess as OP   # noqa: E402
from kiri_ocr import OCR as RefOCR                             # noqa: E402
from kiri_ocr.model import beam_decode_streaming as ref_beam_stream, greedy_decode_streaming as ref_greedy_stream   # noqa: E402
from tests.golden.wide import MARGIN, WIDE_CASES, fit_max_margin, frame_labels, wide_crops   # noqa: E402


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    out = {}
    tmp = tempfile.mkdtemp(prefix="kiri_golden_wide_")
    for name, case in WIDE_CASES.items():
        cfg = CFG()
        sd = FX.make_state_dict(cfg, 202, seed=case["sd_seed"], hardened=True)
        crops, wbs = wide_crops(name)
        # ---------------- CTC head
        planes, mems, feats, labs, seqs = [], [], [], [], []
        for i, (c, Wb) in enumerate(zip(crops, wbs)):
            plane = OP.preprocess_crop(c, 48, Wb)
            x = torch.from_numpy(OP.normalise(plane))[None, None]
            mem = OM.encode(sd, x)
            f = OM._ln(mem, sd, "ctc_head.0")[0]
            lab, seq = frame_labels(name, i, f.shape[0])
            planes.append(plane); mems.append(mem); feats.append(f); labs.append(lab); seqs.append(seq)
        W, b, margin, rn = fit_max_margin(torch.cat(feats), np.concatenate(labs), 204)
        assert margin > 0.999, (name, "ctc fit", margin)
        s = MARGIN / margin
        sd["ctc_head.2.weight"] = (W * s).float().contiguous()
        sd["ctc_head.2.bias"] = (b * s).float().contiguous()
        print(f"{name}: ctc head fitted on {sum(len(l) for l in labs)} frames, margin {margin:.3f} at row norm {rn:.2f} "
              f"(ratio {margin / rn:.3f}); scaled x{s:.2f}")
        # ---------------- decoder head on the teacher-forced states of the same character sequence
        hid, tgt, offs = [], [], []
        for i, mem in enumerate(mems):
            st = OM.DecoderState(sd, OM.mem_proj(sd, mem))
            dec_ids = [c + 1 for c in seqs[i]] + [2]                   # ctc id + 1 = decoder id; EOS
            inp = [1] + dec_ids[:-1]
            for t_in, t_out in zip(inp, dec_ids):
                h = OM.decoder_hidden(st, torch.tensor([t_in]))[0]
                hid.append(h)
                tgt.append(t_out)
                lm = F.linear(h, sd["lm_head.weight"], sd["lm_head.bias"])
                offs.append(cfg.LM_FUSION_ALPHA * F.log_softmax(lm, dim=-1) / MARGIN)
        W, b, margin, rn = fit_max_margin(torch.stack(hid), np.asarray(tgt), 205, offsets=torch.stack(offs).double())
        assert margin > 0.999, (name, "dec fit", margin)
        sd["dec_head.weight"] = (W * MARGIN).float().contiguous()
        sd["dec_head.bias"] = (b * MARGIN).float().contiguous()
        print(f"{name}: dec head fitted on {len(tgt)} steps, margin {margin:.3f} at row norm {rn:.2f}")
        for k, key in (("ctc_w", "ctc_head.2.weight"), ("ctc_b", "ctc_head.2.bias"), ("dec_w", "dec_head.weight"),
                       ("dec_b", "dec_head.bias")):
            out[f"{name}/{k}"] = sd[key].numpy()
        # ---------------- oracle verification + the reference itself
        refs = {}
        for Wb in sorted(set(wbs)):
            rcfg = CFG()
            rcfg.IMG_W = Wb
            path = FX.write_checkpoint(os.path.join(tmp, f"{name}_{Wb}"), sd, rcfg)
            refs[Wb] = {m: RefOCR(model_path=path, device="cpu", decode_method=m) for m in ("fast", "accurate", "beam")}
            assert refs[Wb]["fast"].cfg.IMG_W == Wb
        tok = next(iter(refs.values()))["fast"].tokenizer
        for i, (c, Wb) in enumerate(zip(crops, wbs)):
            key = f"{name}/{i}"
            lg = OM.ctc_logits(sd, mems[i])[0].numpy()
            best, collapsed, conf, length = OD.ctc_greedy(lg)
            srt = np.sort(lg, axis=1)
            ctc_margin = float((srt[:, -1] - srt[:, -2]).min())
            assert np.array_equal(best, labs[i]) and collapsed.tolist() == seqs[i] and length == len(seqs[i])
            assert ctc_margin > 0.9 * MARGIN, (key, ctc_margin)
            want_text = tok.decode_ctc(best.tolist())
            dec_ids = [c + 1 for c in seqs[i]] + [2]
            ocfg = copy.copy(cfg)
            ids, lps, rows = OD.greedy_decode(sd, OM.mem_proj(sd, mems[i]), ocfg, tok.unk_id + 3, length, return_logp=True)
            assert ids == dec_ids, (key, ids, dec_ids)
            top2 = rows.topk(2, dim=1).values
            dec_margin = float((top2[:, 0] - top2[:, 1]).min())
            assert dec_margin > 0.9 * MARGIN, (key, dec_margin)
            # the unmodified reference on the same checkpoint
            pg = np.pad(c, 5, mode="edge")
            box = (5, 5, c.shape[1], c.shape[0])
            r = refs[Wb]
            t = r["fast"]._preprocess_region(pg, box, extra_padding=0)
            plane_ref = ((t[0, 0] * 0.5 + 0.5) * 255.0).round().to(torch.uint8).numpy()
            assert np.array_equal(plane_ref, planes[i]), (key, "plane")
            f_text, f_conf = r["fast"].recognize_region(t)
            a_text, a_conf = r["accurate"].recognize_region(t)
            r["beam"].cfg.BEAM = 3
            b_text, b_conf = r["beam"].recognize_region(t)
            assert f_text == want_text and a_text == want_text and b_text == want_text, (key, f_text, a_text, b_text, want_text)
            o_text, o_conf, _ = OD.recognize_plane(sd, tok, ocfg, planes[i], "decoder")
            assert o_text == a_text and abs(o_conf - a_conf) < 1e-5
            bcfg = copy.copy(cfg)
            bcfg.BEAM = 3
            ob_text, ob_conf, _ = OD.recognize_plane(sd, tok, bcfg, planes[i], "beam")
            assert ob_text == b_text and abs(ob_conf - b_conf) < 1e-5, (key, ob_conf, b_conf)
            with torch.inference_mode():
                m = r["accurate"].model
                mem_r = m.encode(t)
                chunks = list(ref_greedy_stream(m, m.mem_proj(mem_r), tok, r["accurate"].cfg, m.ctc_head(mem_r)))
            stream_ids = [ch["token_id"] for ch in chunks if "token_id" in ch]
            # beam streaming (model.py:949-1152), BEAM 3: per-step best text / confidence / finished, oracle pinned to it
            with torch.inference_mode():
                r["beam"].cfg.BEAM = 3
                bchunks = list(ref_beam_stream(m, m.mem_proj(mem_r), tok, r["beam"].cfg, m.ctc_head(mem_r)))
            ochunks = list(OD.beam_stream_chunks(sd, OM.mem_proj(sd, mems[i]), OM.ctc_logits(sd, mems[i])[0], tok, bcfg))
            assert len(ochunks) == len(bchunks), (key, len(ochunks), len(bchunks))
            for a, b_ in zip(ochunks, bchunks):
                assert a["text"] == b_["text"] and a["token"] == b_["token"] and a["finished"] == b_["finished"] and \
                    abs(a["confidence"] - b_["confidence"]) < 1e-5, (key, a, b_)
            assert bchunks[-1]["finished"] and bchunks[-1]["text"] == want_text
            out[f"{key}/bstream_texts"] = np.array("\x00".join(ch["text"] for ch in bchunks))
            out[f"{key}/bstream_conf"] = np.asarray([ch["confidence"] for ch in bchunks], np.float64)
            out[f"{key}/gstream_conf"] = np.asarray([ch["confidence"] for ch in chunks], np.float64)
            out[f"{key}/gstream_texts"] = np.array("\x00".join(ch["text"] for ch in chunks))
            out[f"{key}/Wb"] = np.int32(Wb)
            out[f"{key}/frame_ids"] = best.astype(np.int16)
            out[f"{key}/ctc_ids"] = collapsed.astype(np.int16)
            out[f"{key}/text"] = np.array(want_text)
            out[f"{key}/fast_conf"] = np.float64(f_conf)
            out[f"{key}/acc_conf"] = np.float64(a_conf)
            out[f"{key}/beam3_conf"] = np.float64(b_conf)
            out[f"{key}/dec_ids"] = np.asarray(ids, np.int16)
            out[f"{key}/stream_ids"] = np.asarray(stream_ids, np.int16)
            out[f"{key}/ctc_margin"] = np.float64(ctc_margin)
            out[f"{key}/dec_margin"] = np.float64(dec_margin)
            print(f"{key}: Wb={Wb} T={len(best)} chars={len(seqs[i])} text={want_text[:24]!r} fast {f_conf:.4f} acc {a_conf:.4f} "
                  f"beam3 {b_conf:.4f} | margins ctc {ctc_margin:.2f} dec {dec_margin:.2f} | stream ids {len(stream_ids)}")
    dst = os.path.join(ROOT, "tests", "golden", "golden_wide_v1.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
