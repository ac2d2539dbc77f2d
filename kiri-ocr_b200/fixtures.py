"""Deterministic synthetic checkpoints, vocabulary and line crops.

There is no network, so no Hugging Face checkpoint, no fonts and no dataset: every test, the
benchmark and the golden vectors use *seeded* random weights in the reference's
``state_dict`` layout (143 tensors) and procedurally drawn line crops
(SURVEY.md §8d).  Everything here is generated with ``numpy.random.default_rng`` (PCG64 —
stable across platforms and numpy versions) so the GPU box regenerates bit-identical
inputs without access to ``/root/reference``.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .config import CFG, META_CONFIG_KEYS


# --------------------------------------------------------------------------- vocabulary
def make_vocab() -> Dict[str, int]:
    """``<unk>`` + printable ASCII (95) + Khmer U+1780..U+17E9 (106) => V = 202."""
    vocab = {"<unk>": 0}
    i = 1
    for cp in list(range(0x20, 0x7F)) + list(range(0x1780, 0x17EA)):
        vocab[chr(cp)] = i
        i += 1
    return vocab


# --------------------------------------------------------------------------- weights
def state_dict_spec(cfg: CFG, vocab_size: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every tensor of ``KiriOCR.state_dict()`` (model.py:235-297)."""
    D, F, Dd, Fd = cfg.ENC_DIM, cfg.ENC_FF, cfg.DEC_DIM, cfg.DEC_FF
    C, Vd = vocab_size + 2, vocab_size + 3
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    chans = [1, 48, 96, 160, D]
    for li, idx in enumerate((0, 3, 6, 9)):
        cin, cout = chans[li], chans[li + 1]
        spec.append((f"stem.net.{idx}.weight", (cout, cin, 3, 3), "conv"))
        b = idx + 1
        spec += [
            (f"stem.net.{b}.weight", (cout,), "bn_gamma"),
            (f"stem.net.{b}.bias", (cout,), "bn_beta"),
            (f"stem.net.{b}.running_mean", (cout,), "bn_mean"),
            (f"stem.net.{b}.running_var", (cout,), "bn_var"),
            (f"stem.net.{b}.num_batches_tracked", (), "counter"),
        ]

    def ln(prefix, d):
        return [(f"{prefix}.weight", (d,), "ln_gamma"), (f"{prefix}.bias", (d,), "ln_beta")]

    def lin(prefix, out_f, in_f, bias=True):
        s = [(f"{prefix}.weight", (out_f, in_f), "linear")]
        if bias:
            s.append((f"{prefix}.bias", (out_f,), "bias"))
        return s

    spec += ln("enc_ln_in", D)
    for l in range(cfg.ENC_LAYERS):
        p = f"enc.layers.{l}"
        spec += [(f"{p}.self_attn.in_proj_weight", (3 * D, D), "linear"),
                 (f"{p}.self_attn.in_proj_bias", (3 * D,), "bias")]
        spec += lin(f"{p}.self_attn.out_proj", D, D)
        spec += lin(f"{p}.linear1", F, D) + lin(f"{p}.linear2", D, F)
        spec += ln(f"{p}.norm1", D) + ln(f"{p}.norm2", D)
    spec += ln("enc_ln", D)
    spec += ln("ctc_head.0", D) + lin("ctc_head.2", C, D)
    spec += lin("mem_proj", Dd, D, bias=False)
    spec.append(("dec_emb.weight", (Vd, Dd), "embedding"))
    spec.append(("dec_pos_enc.pe", (1, cfg.MAX_DEC_LEN + 10, Dd), "pos1d"))
    for l in range(cfg.DEC_LAYERS):
        p = f"dec.layers.{l}"
        for att in ("self_attn", "multihead_attn"):
            spec += [(f"{p}.{att}.in_proj_weight", (3 * Dd, Dd), "linear"),
                     (f"{p}.{att}.in_proj_bias", (3 * Dd,), "bias")]
            spec += lin(f"{p}.{att}.out_proj", Dd, Dd)
        spec += lin(f"{p}.linear1", Fd, Dd) + lin(f"{p}.linear2", Dd, Fd)
        spec += ln(f"{p}.norm1", Dd) + ln(f"{p}.norm2", Dd) + ln(f"{p}.norm3", Dd)
    spec += ln("dec_ln", Dd)
    spec += lin("dec_head", Vd, Dd) + lin("lm_head", Vd, Dd)
    return spec


def sinusoid_1d(length: int, dim: int) -> torch.Tensor:
    """fp32 table built with the same op sequence as model.py:155-161 (bit-identical)."""
    pe = torch.zeros(length, dim)
    position = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def make_state_dict(cfg: Optional[CFG] = None, vocab_size: int = 202, seed: int = 0,
                    hardened: bool = True, head_gain: float = 6.0,
                    eos_bias: float = 0.0, blank_bias: float = 0.0,
                    with_dec_pos_enc: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded random ``state_dict`` in the reference layout.

    ``hardened=False`` imitates torch's default init (BN/LN identity, zero attention biases);
    ``hardened=True`` follows the fixture rules of SURVEY.md §8c: BN running statistics and
    affines, LN affines and every bias are randomised so that folding/bias bugs are visible,
    and the three output heads are scaled by ``head_gain`` to widen top-1 margins.
    ``eos_bias`` / ``blank_bias`` steer the decoder into the EOS-termination branch and the
    CTC head into the all-blank branch (model.py:416-425).
    """
    cfg = cfg or CFG()
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    for key, shape, kind in state_dict_spec(cfg, vocab_size):
        if kind == "conv":
            fan_in = shape[1] * 9
            # stems feed SiLU: He-style gain keeps activations O(1) through four layers
            a = rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        elif kind == "linear":
            bound = 1.0 / math.sqrt(shape[1])
            a = rng.uniform(-bound, bound, shape)
            if "in_proj" in key:
                a = a * math.sqrt(3.0)          # xavier-like (wider than kaiming-uniform)
            if hardened and key.split(".weight")[0] in ("ctc_head.2", "dec_head", "lm_head"):
                a = a * head_gain
        elif kind == "bias":
            if hardened:
                a = rng.normal(0.0, 0.1, shape)
            elif "in_proj" in key or "out_proj" in key:
                a = np.zeros(shape)
            else:
                fan_in = {"linear1": cfg.ENC_DIM, "linear2": cfg.ENC_FF}.get(key.split(".")[-2], cfg.ENC_DIM)
                bound = 1.0 / math.sqrt(fan_in)
                a = rng.uniform(-bound, bound, shape)
        elif kind == "embedding":
            a = rng.standard_normal(shape)
        elif kind in ("bn_gamma", "ln_gamma"):
            a = rng.uniform(0.7, 1.3, shape) if hardened else np.ones(shape)
        elif kind in ("bn_beta", "ln_beta"):
            a = rng.normal(0.0, 0.1, shape) if hardened else np.zeros(shape)
        elif kind == "bn_mean":
            a = rng.normal(0.0, 0.1, shape) if hardened else np.zeros(shape)
        elif kind == "bn_var":
            a = rng.uniform(0.5, 2.0, shape) if hardened else np.ones(shape)
        elif kind == "counter":
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        elif kind == "pos1d":
            if with_dec_pos_enc:
                sd[key] = sinusoid_1d(shape[1], shape[2]).unsqueeze(0)
            continue
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = t(a)

    if eos_bias:
        sd["dec_head.bias"][2] += eos_bias
    if blank_bias:
        sd["ctc_head.2.bias"][0] += blank_bias
    return sd


def write_checkpoint(directory: str, sd: Dict[str, torch.Tensor], cfg: Optional[CFG] = None,
                     name: str = "model") -> str:
    """Write ``<name>.safetensors`` + ``<name>_meta.json`` + ``vocab.json`` exactly as
    ``training.py:1003-1038`` lays them out; returns the safetensors path."""
    from safetensors.torch import save_file

    cfg = cfg or CFG()
    os.makedirs(directory, exist_ok=True)
    vocab_path = os.path.join(directory, "vocab.json")
    with open(vocab_path, "w", encoding="utf-8") as f:
        json.dump(make_vocab(), f, ensure_ascii=False)
    path = os.path.join(directory, f"{name}.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    meta = {"vocab_path": vocab_path, "epoch": 0, "step": 0, "best_val_acc": 0.0,
            "config": {k: getattr(cfg, k) for k in META_CONFIG_KEYS}}
    with open(os.path.join(directory, f"{name}_meta.json"), "w") as f:
        json.dump(meta, f, indent=2)
    return path


# --------------------------------------------------------------------------- line crops
BUCKETS = (128, 256, 384, 512, 640)
BUCKET_SHARE = (0.10, 0.20, 0.25, 0.25, 0.20)


def _draw_line(rng: np.random.Generator, h: int, w: int, inverted: bool) -> np.ndarray:
    """One synthetic text-line crop: light paper (235..255), dark strokes (0..30)."""
    img = rng.integers(235, 256, size=(h, w), dtype=np.int64)
    x = int(rng.integers(2, 8))
    top, bot = max(1, h // 6), max(2, h - h // 6)
    while x < w - 3:
        gw = int(rng.integers(max(2, h // 6), max(3, h // 2) + 1))      # glyph width
        if rng.random() < 0.15:                                            # word gap
            x += gw
            continue
        x1 = min(w - 1, x + gw)
        for _ in range(int(rng.integers(1, 4))):                           # 1-3 strokes / glyph
            ink = int(rng.integers(0, 31))
            if rng.random() < 0.5:                                         # vertical bar
                sx = int(rng.integers(x, x1)) if x1 > x else x
                sw = int(rng.integers(1, max(2, h // 12) + 1))
                y0 = int(rng.integers(top, (top + bot) // 2 + 1))
                y1 = int(rng.integers((top + bot) // 2, bot + 1))
                img[y0:y1, sx:min(w, sx + sw)] = ink
            else:                                                          # horizontal bar
                sy = int(rng.integers(top, bot))
                sh = int(rng.integers(1, max(2, h // 12) + 1))
                img[sy:min(h, sy + sh), x:x1] = ink
        x = x1 + int(rng.integers(1, max(2, h // 10) + 1))
    img = img.astype(np.uint8)
    if inverted:
        img = 255 - img
    return img


def make_line_crops(n: int, seed: int = 1234, bucketed: bool = True,
                    fixed_shape: Optional[Tuple[int, int]] = None) -> List[np.ndarray]:
    """``n`` uint8 crops.  Source heights U{24..96}; widths chosen so that the width after the
    resize to H=48 falls in buckets {128,...,640} with shares {10,20,25,25,20}% (config 2,
    SURVEY §8d); 10 % are dark-background lines to exercise the invert branch (core.py:524)."""
    rng = np.random.default_rng(seed)
    crops = []
    for _ in range(n):
        if fixed_shape is not None:
            h, w = fixed_shape
        else:
            h = int(rng.integers(24, 97))
            b = int(rng.choice(len(BUCKETS), p=BUCKET_SHARE)) if bucketed else len(BUCKETS) - 1
            hi = BUCKETS[b]
            lo = BUCKETS[b - 1] + 1 if b > 0 else 16
            tw = int(rng.integers(lo, hi + 1))                 # width at H=48
            w = max(1, int(round(tw * h / 48.0)))
        crops.append(_draw_line(rng, h, w, inverted=bool(rng.random() < 0.10)))
    return crops


def make_page(n_lines: int = 40, seed: int = 0, page_hw: Tuple[int, int] = (2339, 1654)
              ) -> Tuple[np.ndarray, List[Tuple[int, int, int, int]]]:
    """A synthetic grayscale page with ``n_lines`` text lines and their (x, y, w, h) boxes in
    reading order — the stand-in for a detector's output (config 5)."""
    rng = np.random.default_rng(seed)
    H, W = page_hw
    page = rng.integers(238, 256, size=(H, W), dtype=np.int64).astype(np.uint8)
    boxes = []
    pitch = (H - 80) // n_lines
    for i in range(n_lines):
        lh = int(rng.integers(max(20, pitch // 2), max(21, pitch - 8)))
        lw = int(rng.integers(W // 4, W - 120))
        x = int(rng.integers(40, W - lw - 40))
        y = 40 + i * pitch
        line = _draw_line(rng, lh, lw, inverted=False)
        page[y:y + lh, x:x + lw] = line
        boxes.append((x, y, lw, lh))
    return page, boxes
