"""Oracle: recognizer forward in torch fp32 functional ops.  Test infrastructure.

Restates, straight from a reference-layout ``state_dict`` (no ``nn.Module``):
  * ``ConvStem``                       — kiri_ocr/model.py:211-231 (conv3x3 no-bias, BN eval eps 1e-5, SiLU)
  * ``PosEnc2D`` + H-pool + permute    — kiri_ocr/model.py:176-208, 301-303
  * ``KiriOCR.encode`` LN/encoder/LN   — kiri_ocr/model.py:246-261, 304-306
      (torch ``nn.TransformerEncoderLayer(norm_first=True, activation="gelu")``; third-party
       torch>=2.0.0, pyproject.toml:13 — its published algorithm is restated in ``_enc_layer``)
  * ``ctc_head``                        — kiri_ocr/model.py:263-268
  * ``mem_proj`` + decoder stack        — kiri_ocr/model.py:270-297 (``nn.TransformerDecoderLayer``
      norm_first; cross-attention K/V from the raw projected memory)

All math is fp32 on the CPU under ``torch.inference_mode()``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
LN_EPS = 1e-5
BN_EPS = 1e-5


def _ln(x, sd: SD, prefix: str):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], LN_EPS)


def stem(sd: SD, imgs: torch.Tensor) -> torch.Tensor:
    """[B,1,H,W] -> [B,D,H/8,W/4] (model.py:214-227)."""
    x = imgs
    for idx, stride in ((0, (1, 1)), (3, (2, 2)), (6, (2, 2)), (9, (2, 1))):
        x = F.conv2d(x, sd[f"stem.net.{idx}.weight"], None, stride, 1)
        b = idx + 1
        x = F.batch_norm(x, sd[f"stem.net.{b}.running_mean"], sd[f"stem.net.{b}.running_var"],
                         sd[f"stem.net.{b}.weight"], sd[f"stem.net.{b}.bias"], False, 0.0, BN_EPS)
        x = F.silu(x)
    return x


def sinusoid(length: int, dim: int, dtype=torch.float32) -> torch.Tensor:
    """``PosEnc2D._make_pe`` (model.py:181-192)."""
    pos = torch.arange(length, dtype=dtype).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=dtype) * (-math.log(10000.0) / dim))
    pe = torch.zeros((length, dim), dtype=dtype)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def pos2d_pool(x: torch.Tensor) -> torch.Tensor:
    """PosEnc2D add, mean over H, permute: [B,C,h,w] -> [B,w,C] (model.py:194-208, 302-303)."""
    b, c, h, w = x.shape
    nf = c // 2
    pe_y = sinusoid(h, nf).unsqueeze(2).repeat(1, 1, w)              # [h, nf, w]
    pe_x = sinusoid(w, nf).transpose(0, 1).unsqueeze(0).repeat(h, 1, 1)
    pe = torch.cat([pe_y, pe_x], dim=1).permute(1, 0, 2)              # [2nf, h, w]
    x = x + pe.unsqueeze(0)
    x = x.mean(dim=2)                                                 # adaptive_avg_pool2d -> (1, w)
    return x.permute(0, 2, 1)


def _mha(q_in, k_in, v_in, w, b, wo, bo, heads: int, causal: bool = False):
    """torch ``MultiheadAttention`` forward: packed in_proj rows are [Q; K; V]."""
    D = q_in.shape[-1]
    hd = D // heads
    q = F.linear(q_in, w[:D], b[:D])
    k = F.linear(k_in, w[D:2 * D], b[D:2 * D])
    v = F.linear(v_in, w[2 * D:], b[2 * D:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, heads, hd).transpose(1, 2)
    k = k.view(B, Lk, heads, hd).transpose(1, 2)
    v = v.view(B, Lk, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        mask = torch.triu(torch.ones(Lq, Lk, dtype=torch.bool), diagonal=1 + (Lk - Lq))
        s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, D)
    return F.linear(o, wo, bo)


def _enc_layer(sd: SD, p: str, x, heads: int):
    h = _ln(x, sd, f"{p}.norm1")
    x = x + _mha(h, h, h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"],
                 sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"], heads)
    h = _ln(x, sd, f"{p}.norm2")
    h = F.linear(F.gelu(F.linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                 sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
    return x + h


def n_layers(sd: SD, prefix: str) -> int:
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + ".layers."))


@torch.inference_mode()
def encode(sd: SD, imgs: torch.Tensor, heads: int = 8) -> torch.Tensor:
    """``KiriOCR.encode`` (model.py:299-307): [B,1,48,W] fp32 in [-1,1] -> mem [B,W/4,D]."""
    x = pos2d_pool(stem(sd, imgs))
    x = _ln(x, sd, "enc_ln_in")
    for l in range(n_layers(sd, "enc")):
        x = _enc_layer(sd, f"enc.layers.{l}", x, heads)
    return _ln(x, sd, "enc_ln")


@torch.inference_mode()
def stem_tokens(sd: SD, imgs: torch.Tensor) -> torch.Tensor:
    """Encoder input tokens: LN_in(pos2d_pool(stem)) — a checkpoint for stage parity."""
    return _ln(pos2d_pool(stem(sd, imgs)), sd, "enc_ln_in")


@torch.inference_mode()
def ctc_logits(sd: SD, mem: torch.Tensor) -> torch.Tensor:
    """``ctc_head`` = LN -> Dropout(eval no-op) -> Linear (model.py:264-268)."""
    return F.linear(_ln(mem, sd, "ctc_head.0"), sd["ctc_head.2.weight"], sd["ctc_head.2.bias"])


@torch.inference_mode()
def mem_proj(sd: SD, mem: torch.Tensor) -> torch.Tensor:
    return F.linear(mem, sd["mem_proj.weight"])


# --------------------------------------------------------------------------- decoder
class DecoderState:
    """Per-line KV cache for single-token decoder steps (result-preserving restatement of the
    reference's full-prefix re-run, model.py:465-479; verified in tests/golden/make_golden.py)."""

    def __init__(self, sd: SD, memp: torch.Tensor, heads: int = 8):
        self.sd, self.heads = sd, heads
        self.L = n_layers(sd, "dec")
        D = memp.shape[-1]
        self.cross_k, self.cross_v = [], []
        for l in range(self.L):
            w = sd[f"dec.layers.{l}.multihead_attn.in_proj_weight"]
            b = sd[f"dec.layers.{l}.multihead_attn.in_proj_bias"]
            self.cross_k.append(F.linear(memp, w[D:2 * D], b[D:2 * D]))
            self.cross_v.append(F.linear(memp, w[2 * D:], b[2 * D:]))
        self.self_k: List[Optional[torch.Tensor]] = [None] * self.L
        self.self_v: List[Optional[torch.Tensor]] = [None] * self.L
        self.pos = 0


def _attend(q, k, v, heads: int):
    B, Lq, D = q.shape
    hd = D // heads
    Lk = k.shape[1]
    qh = q.view(B, Lq, heads, hd).transpose(1, 2)
    kh = k.view(B, Lk, heads, hd).transpose(1, 2)
    vh = v.view(B, Lk, heads, hd).transpose(1, 2)
    p = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, D)


@torch.inference_mode()
def decoder_hidden(st: DecoderState, token_ids: torch.Tensor) -> torch.Tensor:
    """One token per line through the decoder stack: returns ``dec_ln`` output [B, D] (the input of
    ``dec_head`` / ``lm_head``) and advances the KV cache.

    Input embedding is ``dec_emb[id] + pe[pos]`` with no sqrt(d) scale (model.py:465-470); old
    checkpoints without ``dec_pos_enc.pe`` skip the positional term (core.py:255-263)."""
    sd, heads = st.sd, st.heads
    x = sd["dec_emb.weight"][token_ids].unsqueeze(1)                  # [B,1,D]
    if "dec_pos_enc.pe" in sd:
        x = x + sd["dec_pos_enc.pe"][:, st.pos:st.pos + 1, :]
    D = x.shape[-1]
    for l in range(st.L):
        p = f"dec.layers.{l}"
        h = _ln(x, sd, f"{p}.norm1")
        w, b = sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"]
        q = F.linear(h, w[:D], b[:D])
        k = F.linear(h, w[D:2 * D], b[D:2 * D])
        v = F.linear(h, w[2 * D:], b[2 * D:])
        st.self_k[l] = k if st.self_k[l] is None else torch.cat([st.self_k[l], k], dim=1)
        st.self_v[l] = v if st.self_v[l] is None else torch.cat([st.self_v[l], v], dim=1)
        a = _attend(q, st.self_k[l], st.self_v[l], heads)
        x = x + F.linear(a, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        h = _ln(x, sd, f"{p}.norm2")
        w, b = sd[f"{p}.multihead_attn.in_proj_weight"], sd[f"{p}.multihead_attn.in_proj_bias"]
        q = F.linear(h, w[:D], b[:D])
        a = _attend(q, st.cross_k[l], st.cross_v[l], heads)
        x = x + F.linear(a, sd[f"{p}.multihead_attn.out_proj.weight"], sd[f"{p}.multihead_attn.out_proj.bias"])
        h = _ln(x, sd, f"{p}.norm3")
        h = F.linear(F.gelu(F.linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                     sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
        x = x + h
    st.pos += 1
    return _ln(x, sd, "dec_ln")[:, 0, :]


@torch.inference_mode()
def decoder_step(st: DecoderState, token_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """One token per line: returns (dec_head logits, lm_head logits), each [B, Vd] (model.py:478-484)."""
    sd = st.sd
    out = decoder_hidden(st, token_ids)
    dec = F.linear(out, sd["dec_head.weight"], sd["dec_head.bias"])
    lm = F.linear(out, sd["lm_head.weight"], sd["lm_head.bias"]) if "lm_head.weight" in sd else None
    return dec, lm
