"""``OCR`` — the reference's document API with the recognition stage running on the B200 engine.

Drop-in for ``kiri_ocr.core.OCR`` (kiri_ocr/core.py:40-1161): same constructor arguments,
decode-method aliases and errors (core.py:56-156), same result/chunk schemas
(core.py:778-784, 859-866, 963-1000), same checkpoint files (core.py:219-465), same detector
plug-in boundary (core.py:469-485, 746-756: anything with ``detect_lines_objects`` /
``detect_lines`` / ``detect_words``).  What changes is the loop body: instead of one
``_preprocess_region`` + ``recognize_region`` per box on the host, all boxes of a page go through
``BatchedRecognizer`` in one batch.  There is no CPU path: ``device`` must be a CUDA device.
``decode_method="beam"`` runs the device beam search (widths 1..5, ``ocr.cfg.BEAM``) and the device CTC
forward rescoring.  The streaming calls are LIVE for the decoders: the persistent decode kernel publishes every
step into mapped host memory and chunks are yielded while it is still running (beam streaming follows
``beam_decode_streaming``'s own rule: prune by ``score / L**0.8``, stop when the best hypothesis ended).
"""
from __future__ import annotations

import json
import sys
import warnings
from pathlib import Path
from typing import Dict, Generator, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .config import CFG, CharTokenizer
from .engine import BatchedRecognizer, LineResult

DECODE_ALIASES = {"fast": "ctc", "ctc": "ctc", "accurate": "decoder", "decoder": "decoder", "beam": "beam"}


class OCR:
    _model_cache: Dict[Tuple[str, str], Dict] = {}

    def __init__(self, model_path: str = "mrrtmob/kiri-ocr", det_model_path: Optional[str] = None,
                 det_method: str = "db", det_conf_threshold: float = 0.5, padding: int = 10,
                 device: str = "cuda", verbose: bool = False, decode_method: str = "accurate",
                 use_beam_search: Optional[bool] = None, use_fp16: Optional[bool] = None,
                 width_mode: str = "parity"):
        if use_beam_search is not None:
            warnings.warn("use_beam_search is deprecated. Use decode_method instead:\n"
                          "  - decode_method='fast' (replaces use_beam_search=False)\n"
                          "  - decode_method='accurate' (default, balanced)\n"
                          "  - decode_method='beam' (replaces use_beam_search=True)",
                          DeprecationWarning, stacklevel=2)
            decode_method = "beam" if use_beam_search else "fast"
        decode_method = self._normalize_decode_method(decode_method)
        if not str(device).startswith("cuda"):
            raise _lib.KiriError(f"kiri_ocr_b200.OCR runs the recognizer on B200 CUDA kernels only; device={device!r} "
                                 "has no implementation here (use the reference package for CPU)")
        self.device = device
        self.verbose = verbose
        self.padding = padding
        self.det_model_path = det_model_path
        self.det_method = det_method
        self.det_conf_threshold = det_conf_threshold
        self.decode_method = decode_method
        self.use_fp16 = use_fp16
        self.use_beam_search = decode_method == "beam"
        self.width_mode = width_mode
        self.cfg: Optional[CFG] = None
        self.tokenizer: Optional[CharTokenizer] = None
        self.model: Optional[BatchedRecognizer] = None
        self.repo_id: Optional[str] = None
        if ("/" in model_path and not model_path.startswith((".", "/"))
                and not model_path.endswith((".safetensors", ".pt", ".onnx", ".pth"))):
            self.repo_id = model_path
        self._load_model(self._resolve_model_path(model_path))
        self._detector = None

    @staticmethod
    def _normalize_decode_method(method: str) -> str:
        method = method.lower().strip()
        if method not in DECODE_ALIASES:
            raise ValueError(f"Invalid decode_method '{method}'. "
                             f"Choose from: 'fast', 'accurate', 'beam' (or aliases: 'ctc', 'decoder')")
        return DECODE_ALIASES[method]

    # ==================== model loading (core.py:160-465) ====================
    def _resolve_model_path(self, model_path: str) -> str:
        f = Path(model_path)
        if f.exists():
            return str(f)
        pkg = Path(__file__).parent
        for cand in (pkg / model_path, pkg.parent / "models" / f.name):
            if cand.exists():
                return str(cand)
        if "/" in model_path and not model_path.startswith((".", "/")):
            return self._download_from_huggingface(model_path)
        return model_path

    def _download_from_huggingface(self, repo_id: str) -> str:
        try:
            from huggingface_hub import hf_hub_download
            for name in ("config.json", "vocab.json", "vocab_auto.json"):
                try:
                    hf_hub_download(repo_id=repo_id, filename=name)
                except Exception:
                    pass
            for name in ("model.safetensors", "model.pt"):
                try:
                    return hf_hub_download(repo_id=repo_id, filename=name)
                except Exception:
                    pass
        except Exception as e:                                    # no network / no hub package
            if self.verbose:
                print(f"HuggingFace download failed: {e}")
        return repo_id

    def _load_model(self, model_path: str) -> None:
        key = (str(model_path), f"{self.device}|{self.width_mode}")
        if key in OCR._model_cache:
            c = OCR._model_cache[key]
            self.model, self.cfg, self.tokenizer = c["model"], c["cfg"], c["tokenizer"]
            return
        try:
            if model_path.endswith(".safetensors"):
                sd, vocab_path = self._load_safetensors(model_path)
            else:
                sd, vocab_path = self._load_torch_checkpoint(model_path)
            if self.use_fp16 is not None:
                self.cfg.USE_FP16 = self.use_fp16
            vocab_path = self._find_vocab_file(vocab_path, model_path)
            if not vocab_path or not Path(vocab_path).exists():
                raise FileNotFoundError(f"Could not find vocabulary file. Expected near: {model_path}")
            self.tokenizer = CharTokenizer(vocab_path, self.cfg)
            want = {"ctc_head.2.weight": self.tokenizer.ctc_classes, "dec_head.weight": self.tokenizer.dec_vocab}
            for k, rows in want.items():
                if k in sd and sd[k].shape[0] != rows:
                    raise RuntimeError(f"size mismatch for {k}: checkpoint has {sd[k].shape[0]} rows, vocabulary needs {rows}")
            self.model = BatchedRecognizer(sd, self.cfg, self.tokenizer, device=self.device, width_mode=self.width_mode)
            OCR._model_cache[key] = {"model": self.model, "cfg": self.cfg, "tokenizer": self.tokenizer}
        except RuntimeError as e:
            if "size mismatch" in str(e):
                print(f"\nModel/vocab size mismatch: {e}")
                sys.exit(1)
            raise

    def _load_safetensors(self, model_path: str):
        from safetensors.torch import load_file
        sd = load_file(model_path, device="cpu")
        self.cfg = CFG()
        vocab_path = ""
        meta = model_path.replace(".safetensors", "_meta.json")
        if Path(meta).exists():
            with open(meta, "r") as f:
                md = json.load(f)
            vocab_path = md.get("vocab_path", "")
            self._apply_config(md.get("config", {}))
        else:
            self._infer_config_from_state_dict(sd)
        return sd, vocab_path

    def _load_torch_checkpoint(self, model_path: str):
        ck = torch.load(model_path, map_location="cpu", weights_only=False)
        if "config" in ck:
            cd = ck["config"]
            self.cfg = CFG()
            if isinstance(cd, dict):
                self._apply_config(cd)
            else:                                                  # a pickled CFG-like object
                for k in vars(self.cfg):
                    if hasattr(cd, k):
                        setattr(self.cfg, k, getattr(cd, k))
            return ck["model"], ck.get("vocab_path", "")
        self.cfg = CFG()
        return ck, ""

    def _infer_config_from_state_dict(self, sd: Dict) -> None:
        """Fallback for checkpoints without metadata (core.py:319-403)."""
        c = self.cfg
        if "stem.net.9.weight" in sd:
            c.ENC_DIM = sd["stem.net.9.weight"].shape[0]
        for prefix, attr in (("enc.layers.", "ENC_LAYERS"), ("dec.layers.", "DEC_LAYERS")):
            n = {int(k.split(".")[2]) for k in sd if k.startswith(prefix)}
            if n:
                setattr(c, attr, max(n) + 1)
        if "enc.layers.0.linear1.weight" in sd:
            c.ENC_FF = sd["enc.layers.0.linear1.weight"].shape[0]
        if "dec_emb.weight" in sd:
            c.DEC_DIM = sd["dec_emb.weight"].shape[1]
        if "dec.layers.0.linear1.weight" in sd:
            c.DEC_FF = sd["dec.layers.0.linear1.weight"].shape[0]
        for key, attr in (("enc.layers.0.self_attn.in_proj_weight", "ENC_HEADS"),
                          ("dec.layers.0.self_attn.in_proj_weight", "DEC_HEADS")):
            if key in sd:
                d = sd[key].shape[0] // 3
                setattr(c, attr, d // 64 if d % 64 == 0 else (d // 32 if d % 32 == 0 else 8))

    def _apply_config(self, cd: Dict) -> None:
        if not cd:
            return
        for k in ("IMG_H", "IMG_W", "ENC_DIM", "ENC_LAYERS", "ENC_HEADS", "ENC_FF", "DEC_DIM", "DEC_LAYERS",
                  "DEC_HEADS", "DEC_FF", "DROPOUT", "USE_CTC", "USE_FP16"):
            setattr(self.cfg, k, cd.get(k, getattr(self.cfg, k)))

    def _find_vocab_file(self, vocab_path: str, model_path: str) -> Optional[str]:
        d = Path(model_path).parent
        for cand in (vocab_path, d / Path(vocab_path).name if vocab_path else None, d / "vocab.json",
                     d / "vocab_auto.json", d / "vocab_char.json"):
            if cand and Path(cand).exists():
                return str(cand)
        return None

    # ==================== detector plug-in (core.py:469-485) ====================
    @property
    def detector(self):
        if self._detector is None:
            try:
                from kiri_ocr.detector import TextDetector          # the reference's own detectors
            except Exception as e:
                raise _lib.KiriError("no text detector available: install kiri-ocr for its TextDetector or set "
                                     "`ocr._detector` to any object with detect_lines_objects()/detect_lines()/"
                                     f"detect_words() ({e})")
            det_path = self.det_model_path
            if det_path is None and self.repo_id and self.det_method in ("db", "craft"):
                det_path = self.repo_id
            self._detector = TextDetector(method=self.det_method, model_path=det_path,
                                          conf_threshold=self.det_conf_threshold)
        return self._detector

    def _detect(self, image_path, mode: str):
        if mode == "lines":
            if hasattr(self.detector, "detect_lines_objects"):
                tb = self.detector.detect_lines_objects(image_path)
                return [b.bbox for b in tb], [b.confidence for b in tb]
            boxes = self.detector.detect_lines(image_path)
        else:
            boxes = self.detector.detect_words(image_path)
        return boxes, [1.0] * len(boxes)

    @staticmethod
    def _read_image(image_path) -> np.ndarray:
        """cv2.imread as the reference does (core.py:762-764): BGR uint8 [H, W, 3] (or gray [H, W])."""
        import cv2
        img = cv2.imread(str(image_path))
        if img is None:
            raise ValueError(f"Could not load image: {image_path}")
        return img

    @classmethod
    def _read_gray(cls, image_path) -> np.ndarray:
        import cv2
        img = cls._read_image(image_path)
        return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if len(img.shape) == 3 else img

    # ==================== recognition ====================
    def _method(self, decode_method: Optional[str] = None) -> str:
        m = self._normalize_decode_method(decode_method) if decode_method is not None else self.decode_method
        return m

    def _preprocess_region(self, img: np.ndarray, box, extra_padding: int = 5) -> Optional[torch.Tensor]:
        """Same contract as core.py:489-528 — fp32 ``[1,1,IMG_H,IMG_W]`` in [-1,1] or None — computed by
        the device kernel (bit-identical planes)."""
        if img.ndim == 3:
            import cv2
            img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        eng = self.model
        ent, valid = eng.boxes_to_entries(img.shape[:2], [box], extra_padding=extra_padding)
        if not valid[0]:
            return None
        page = np.ascontiguousarray(img)
        from .engine import plan_groups
        (idx, descs, smem, n_strips), = plan_groups(ent, self.cfg, "parity").values()
        with torch.cuda.stream(eng.stream):
            src, _ = eng._stage_host([page], 0)                  # persistent pinned staging, no per-call cudaHostAlloc
            planes, _ = eng.preprocess(src.to(eng.device, non_blocking=True), descs, self.cfg.IMG_W, smem, n_strips)
            t = planes[0].float().cpu() / 255.0                  # (synchronises: the staging buffer is free again)
        return ((t - 0.5) / 0.5).unsqueeze(0).unsqueeze(0)

    @staticmethod
    def _tensor_to_plane(image_tensor: torch.Tensor) -> torch.Tensor:
        t = image_tensor.detach().float().cpu().reshape(image_tensor.shape[-2], image_tensor.shape[-1])
        return torch.round((t * 0.5 + 0.5) * 255.0).clamp(0, 255).to(torch.uint8)

    def _recognize_planes(self, planes: torch.Tensor, method: str, streaming: bool = False) -> List[LineResult]:
        """Already preprocessed uint8 planes [n, IMG_H, W]: they pass through the resample kernel as the identity and
        are NEVER inverted (OCR.recognize_region takes the tensor as it is, core.py:530-568; the dark-background
        test belongs to ``_preprocess_region`` only)."""
        eng = self.model
        n, H, W = planes.shape
        ent = np.array([(i * H * W, W, W, H, _lib.CROP_NO_INVERT) for i in range(n)], np.int64)
        return eng.recognize_packed([planes.numpy().reshape(-1)], ent, method, streaming)

    def recognize_region(self, image_tensor: torch.Tensor) -> Tuple[str, float]:
        r = self._recognize_planes(self._tensor_to_plane(image_tensor)[None], self._method())[0]
        return r.text, r.confidence

    def recognize_region_streaming(self, image_tensor: torch.Tensor, decode_method: Optional[str] = None
                                   ) -> Generator[Dict, None, None]:
        yield from self._stream_planes(self._tensor_to_plane(image_tensor)[None], self._method(decode_method))

    def _chunks(self, r: LineResult, method: str) -> Generator[Dict, None, None]:
        """Replay a decoded line as the reference's streaming chunks (model.py:736-775, 845-946)."""
        tok = self.tokenizer
        if method == "ctc":
            text, prev, step = "", None, 0
            for idx, p in zip(r.frame_ids.tolist(), r.frame_prob.tolist()):
                if idx == prev:
                    continue
                prev = idx
                if idx < tok.ctc_offset:
                    continue
                raw = idx - tok.ctc_offset
                if 0 <= raw < tok.vocab_size:
                    ch = tok.id_to_token.get(raw, "")
                    if ch and ch != tok.unk_token:
                        text += ch
                        step += 1
                        yield {"token": ch, "token_id": idx, "text": text, "confidence": float(p), "step": step,
                               "finished": False}
            yield {"token": "", "token_id": -1, "text": text, "confidence": float(r.ctc_confidence), "step": step,
                   "finished": True}
            return
        text = ""
        for step, (tid, p) in enumerate(zip(r.ids.tolist(), r.step_prob.tolist())):
            finished = tid == tok.dec_eos
            ch = ""
            if not finished and tid not in (tok.dec_pad, tok.dec_bos, tok.dec_eos):
                raw = tid - tok.dec_offset
                if 0 <= raw < tok.vocab_size:
                    ch = tok.id_to_token.get(raw, "")
                    if ch != tok.unk_token:
                        text += ch
            yield {"token": ch, "token_id": tid, "text": text, "confidence": float(p), "step": step + 1,
                   "finished": finished}
            if finished:
                break

    def _live_chunks(self, tk, line: int, method: str) -> Generator[Dict, None, None]:
        """The reference's streaming chunks for one line of a LIVE ticket, produced while the device is still decoding
        (greedy_decode_streaming model.py:845-946, beam_decode_streaming model.py:1117-1150)."""
        tok, eng = self.tokenizer, self.model
        if method == "decoder":
            text = ""
            for tid, prob, _, step, finished in eng.live_greedy(tk, line):
                ch = ""
                if not finished and tid not in (tok.dec_pad, tok.dec_bos, tok.dec_eos) and 0 <= tid < tok.dec_vocab:
                    ch = tok.dec_text[tid]                       # '' for <unk> (model.py:927-929)
                    text += ch
                yield {"token": ch, "token_id": tid, "text": text, "confidence": prob, "step": step, "finished": finished}
            return
        import math
        prev = ""
        for ids, lps, step, finished in eng.live_beam(tk, line):
            cut = ids[: ids.index(tok.dec_eos)] if tok.dec_eos in ids else ids
            cur = tok.decode_dec(cut)
            conf = min(1.0, max(0.0, math.exp(sum(lps) / len(lps)))) if lps else 0.0
            yield {"token": cur[len(prev):] if len(cur) > len(prev) else "", "text": cur, "confidence": conf, "step": step,
                   "finished": finished}
            prev = cur

    def _stream_planes(self, planes: torch.Tensor, method: str) -> Generator[Dict, None, None]:
        """Character stream of ONE preprocessed plane: CTC replays the per-frame decisions of the fused kernel
        (the whole line is one launch); the decoders stream live from the persistent kernel."""
        eng = self.model
        n, H, W = planes.shape
        ent = np.array([(i * H * W, W, W, H, _lib.CROP_NO_INVERT) for i in range(n)], np.int64)
        if method == "ctc":
            yield from self._chunks(eng.recognize_packed([planes.numpy().reshape(-1)], ent, "ctc", True)[0], "ctc")
            return
        with torch.cuda.stream(eng.stream):
            tk = eng.submit([planes.numpy().reshape(-1)], ent, method, streaming=True, live=True)
        try:
            yield from self._live_chunks(tk, 0, method)
        finally:
            eng.live_finish(tk)

    def _single_line_plane(self, image_path) -> torch.Tensor:
        img = self._read_gray(image_path)
        t = self._preprocess_region(img, (0, 0, img.shape[1], img.shape[0]), extra_padding=0)
        return self._tensor_to_plane(t)

    def recognize_single_line_image(self, image_path: Union[str, Path]) -> Tuple[str, float]:
        r = self._recognize_planes(self._single_line_plane(image_path)[None], self._method())[0]
        return r.text, r.confidence

    def recognize_streaming(self, image_path: Union[str, Path], decode_method: Optional[str] = None
                            ) -> Generator[Dict, None, None]:
        yield from self._stream_planes(self._single_line_plane(image_path)[None], self._method(decode_method))

    # ==================== documents ====================
    def _recognize_document(self, image_path, mode: str, method: str, streaming: bool = False):
        boxes, det_confs = self._detect(image_path, mode)
        # the page travels to the device as cv2 decoded it; BGR -> gray (core.py:766) runs there, bit-exact with cv2
        img = self._read_image(image_path)
        res = self.model.recognize_boxes(img, boxes, method, streaming) if len(boxes) else []
        return boxes, det_confs, res

    @staticmethod
    def _box_list(box):
        try:
            return [int(v) for v in box]
        except Exception:                                          # noqa: BLE001 - a malformed detector box
            return list(box) if isinstance(box, (list, tuple)) else [box]

    def _results(self, boxes, det_confs, res, total: Optional[int] = None, keep_errors: bool = False):
        """Per-region result dicts (core.py:778-784).  Empty crops are skipped (core.py:773-774); a region that
        failed is dropped like the reference's swallowed per-region exception (core.py:789-791) or, for the
        streaming form, reported with an ``error`` key (core.py:873-885)."""
        from .engine import LineError
        for i, (box, dc, r) in enumerate(zip(boxes, det_confs, res), 1):
            if r is None:
                continue
            if isinstance(r, LineError):
                if keep_errors:
                    yield {"box": self._box_list(box), "text": "", "confidence": 0.0, "det_confidence": float(dc),
                           "line_number": i, "total_regions": total, "error": r.message}
                elif self.verbose:
                    print(f"  {i:2d}. [Error: {r.message}]")
                continue
            d = {"box": [int(v) for v in box], "text": r.text, "confidence": float(r.confidence),
                 "det_confidence": float(dc), "line_number": i}
            if total is not None:
                d["total_regions"] = total
            yield d

    def process_document(self, image_path: Union[str, Path], mode: str = "lines", verbose: bool = False) -> List[Dict]:
        boxes, det_confs, res = self._recognize_document(image_path, mode, self._method())
        out = list(self._results(boxes, det_confs, res))
        if verbose:
            for d in out:
                print(f"  {d['line_number']:2d}. {d['text'][:50]:50s} ({d['confidence'] * 100:.1f}%)")
        return out

    def process_documents(self, image_paths: List[Union[str, Path]], mode: str = "lines") -> List[List[Dict]]:
        """Several pages at once (an extension; the reference has only the per-page call): detection and image
        decoding stay per page on the host, recognition runs through ``recognize_pages`` — whole pages per batch,
        each uploaded once, two batches in flight.  Returns ``process_document``'s list for every page."""
        method = self._method()
        det, pages = [], []
        for path in image_paths:
            det.append(self._detect(path, mode))
            pages.append(self._read_image(path))
        res = self.model.recognize_pages(pages, [d[0] for d in det], method)
        return [list(self._results(b, c, r)) for (b, c), r in zip(det, res)]

    def process_document_streaming(self, image_path: Union[str, Path], mode: str = "lines", verbose: bool = False
                                   ) -> Generator[Dict, None, None]:
        boxes, det_confs, res = self._recognize_document(image_path, mode, self._method())
        yield from self._results(boxes, det_confs, res, total=len(boxes), keep_errors=True)

    def extract_text_stream_chars(self, image_path: Union[str, Path], mode: str = "lines",
                                  decode_method: Optional[str] = None, verbose: bool = False
                                  ) -> Generator[Dict, None, None]:
        from .engine import LineError
        method = self._method(decode_method)
        eng = self.model
        boxes, det_confs = self._detect(image_path, mode)
        img = self._read_image(image_path)
        total = len(boxes)
        tk, res = None, []
        if total:
            if method == "ctc":
                res = eng.recognize_boxes(img, boxes, "ctc", True)
            else:
                # LIVE: all regions of the page are submitted as one batch; the decode kernel publishes every step into
                # mapped host memory, and region k's characters are yielded while regions k+1.. are still decoding
                ent, idx, errors = eng._page_entries(img.shape[:2], boxes, 0)
                res = [None] * total
                for i, e in errors.items():
                    res[i] = e
                if len(ent):
                    bgr = [(0, 0, img.shape[0] * img.shape[1], True)] if img.ndim == 3 else None
                    with torch.cuda.stream(eng.stream):
                        tk = eng.submit([np.ascontiguousarray(img)], ent, method, streaming=True, bgr_pages=bgr, live=True)
                    for k, i in enumerate(idx):
                        res[int(i)] = k                          # line index inside the ticket
        done: List[str] = []
        try:
            for num, (box, dc, r) in enumerate(zip(boxes, det_confs, res), 1):
                if r is None:
                    continue
                if isinstance(r, LineError):                       # core.py:1011-1026
                    yield {"token": "", "text": "", "cumulative_text": "\n".join(done), "region_number": num,
                           "total_regions": total, "step": 0, "region_finished": True, "document_finished": num == total,
                           "region_start": True, "box": self._box_list(box), "error": r.message}
                    continue
                b = [int(v) for v in box]
                yield {"token": "", "text": "", "cumulative_text": "\n".join(done), "region_number": num,
                       "total_regions": total, "step": 0, "region_finished": False, "document_finished": False,
                       "region_start": True, "box": b, "det_confidence": float(dc)}
                cur = ""
                chunks = self._chunks(r, method) if method == "ctc" else self._live_chunks(tk, r, method)
                for ch in chunks:
                    cur = ch["text"]
                    yield {"token": ch["token"], "text": cur, "cumulative_text": "\n".join(done + ([cur] if cur else [])),
                           "region_number": num, "total_regions": total, "step": ch["step"],
                           "confidence": ch["confidence"], "region_finished": ch["finished"],
                           "document_finished": ch["finished"] and num == total, "region_start": False, "box": b,
                           "det_confidence": float(dc)}
                    if ch["finished"]:
                        break
                if cur:
                    done.append(cur)
        finally:
            if tk is not None:
                eng.live_finish(tk)

    @staticmethod
    def _group_lines(results: List[Dict]) -> List[str]:
        """Reading-order grouping by 0.8 * max-height tolerance (core.py:1128-1160)."""
        lines, cur, prev_c, prev_h = [], [], None, None
        for res in results:
            y, h = res["box"][1], res["box"][3]
            c = y + h / 2
            if prev_c is not None and abs(c - prev_c) < max(h, prev_h) * 0.8:
                cur.append(res["text"])
            else:
                if prev_c is not None:
                    lines.append(" ".join(cur))
                cur = [res["text"]]
            prev_c, prev_h = c, h
        if cur:
            lines.append(" ".join(cur))
        return lines

    def extract_text_streaming(self, image_path: Union[str, Path], mode: str = "lines", verbose: bool = False
                               ) -> Generator[Dict, None, None]:
        seen: List[Dict] = []
        for result in self.process_document_streaming(image_path, mode, verbose):
            if "error" not in result and result["text"]:
                seen.append(result)
            result["cumulative_text"] = "\n".join(self._group_lines(seen))
            yield result

    def extract_text(self, image_path: Union[str, Path], mode: str = "lines", verbose: bool = False
                     ) -> Tuple[str, List[Dict]]:
        results = self.process_document(image_path, mode, verbose=verbose)
        if not results:
            return "", results
        return "\n".join(self._group_lines(results)), results
