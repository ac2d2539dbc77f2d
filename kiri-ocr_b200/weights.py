"""One-time packing of a reference ``state_dict`` into the layouts the kernels read.

Consumes exactly the tensors ``KiriOCR.state_dict()`` holds (kiri_ocr/model.py:235-297; layout
listed in SURVEY.md §8b) and produces, on the target device:

* conv stem: eval-mode BatchNorm folded into the conv weights (``w' = w*g/sqrt(var+1e-5)``,
  ``b' = beta - mean*g/sqrt(var+1e-5)``), laid out ``[Cout, (ky, kx, cin)]`` in bf16 with ``cin``
  padded to a multiple of 32 (the K-chunk of the tcgen05 kernel); layer 1 stays fp32 on the
  host because it travels as a kernel parameter;
* the constant 2-D positional table ``mean_h(pe_y) | pe_x`` (model.py:194-208 folded through the
  H-pool);
* Linear weights ``[out, in]`` in bf16, biases and LayerNorm affines in fp32; the CTC head and
  the two decoder heads zero-padded to a multiple of 16 rows;
* cross-attention K/V projection fused with ``mem_proj``: ``(W_k; W_v)_l @ W_memproj`` so the
  per-batch cross-K/V precompute is one GEMM straight from the encoder memory.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List

import torch

from . import _lib
from .config import CFG

BN_EPS = 1e-5


def _sinusoid(length: int, dim: int) -> torch.Tensor:
    # same op order as PosEnc2D._make_pe (model.py:181-192), fp32
    pos = torch.arange(length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros((length, dim), dtype=torch.float32)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def pos_table(rows: int, max_t: int, dim: int) -> torch.Tensor:
    nf = dim // 2
    pe_y = _sinusoid(rows, nf).mean(dim=0, keepdim=True).expand(max_t, nf)
    pe_x = _sinusoid(max_t, nf)
    return torch.cat([pe_y, pe_x], dim=1).contiguous()


def fold_bn(sd: Dict[str, torch.Tensor], conv_idx: int):
    w = sd[f"stem.net.{conv_idx}.weight"].float()
    b = conv_idx + 1
    g, beta = sd[f"stem.net.{b}.weight"].float(), sd[f"stem.net.{b}.bias"].float()
    mu, var = sd[f"stem.net.{b}.running_mean"].float(), sd[f"stem.net.{b}.running_var"].float()
    s = g / torch.sqrt(var + BN_EPS)
    return w * s[:, None, None, None], beta - mu * s


def conv_gemm_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9 * Cin_pad] ordered (ky, kx, cin), cin padded to 32."""
    cout, cin = w.shape[:2]
    cin_pad = (cin + 31) // 32 * 32
    k = torch.zeros(cout, 3, 3, cin_pad, dtype=torch.float32)
    k[..., :cin] = w.permute(0, 2, 3, 1)
    return k.reshape(cout, 9 * cin_pad)


class PackedWeights:
    """Device tensors + the ``KiriWeights`` / ``KiriDims`` structs that point at them."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: CFG, vocab_size: int, device: torch.device):
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        self.cfg = cfg
        self.device = device
        self._keep: List[torch.Tensor] = []
        D, Dd = cfg.ENC_DIM, cfg.DEC_DIM
        self.C = vocab_size + 2
        self.Vd = vocab_size + 3
        self.Cp = (self.C + 15) // 16 * 16
        self.Vp = (self.Vd + 15) // 16 * 16
        enc_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("enc.layers."))
        dec_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("dec.layers."))
        self.has_dec_pos = "dec_pos_enc.pe" in sd
        self.has_lm = "lm_head.weight" in sd
        max_t = cfg.IMG_W // 4
        max_pos = cfg.MAX_DEC_LEN + 10

        W = _lib.KiriWeights()
        dev_bf16 = lambda t: self._dev(t.to(torch.bfloat16))      # noqa: E731
        dev_f32 = lambda t: self._dev(t.to(torch.float32))        # noqa: E731

        w1, b1 = fold_bn(sd, 0)
        self.conv1_w = w1.reshape(48, 9).contiguous()
        self.conv1_b = b1.contiguous()
        W.conv1_w_host, W.conv1_b_host = self.conv1_w.data_ptr(), self.conv1_b.data_ptr()
        for name, idx in (("conv2", 3), ("conv3", 6), ("conv4", 9)):
            w, b = fold_bn(sd, idx)
            setattr(W, f"{name}_w", dev_bf16(conv_gemm_weight(w)))
            setattr(W, f"{name}_b", dev_f32(b))
        W.pos_table = dev_f32(pos_table(cfg.IMG_H // 8, max_t, D))
        W.enc_ln_in_g, W.enc_ln_in_b = dev_f32(sd["enc_ln_in.weight"]), dev_f32(sd["enc_ln_in.bias"])
        # The affine of norm2 is folded into linear1 and, from layer 1 on, the affine of norm1 into the QKV projection
        # (W' = W diag(g), b' = b + W beta, in fp64 from the fp32 checkpoint, rounded to bf16 once): the encoder-tail
        # kernel's two LayerNorms are then pure normalisations and the library is handed identity affines for them
        # (csrc/encoder_block.cu picks its parameter-free epilogue when it sees gain 1 / shift 0).  Layer 0's norm1 is
        # applied by the pool + LayerNorm kernel and keeps its affine.
        ones, zeros = torch.ones(D), torch.zeros(D)

        def folded(w, b, g, beta):
            w64 = w.double()
            return (w64 * g.double()[None, :]).float(), (b.double() + w64 @ beta.double()).float()

        for l in range(enc_layers):
            p, L = f"enc.layers.{l}", W.enc[l]
            wqkv, bqkv = sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"]
            if l > 0:
                wqkv, bqkv = folded(wqkv, bqkv, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"])
                L.ln1_g, L.ln1_b = dev_f32(ones), dev_f32(zeros)
            else:
                L.ln1_g, L.ln1_b = dev_f32(sd[f"{p}.norm1.weight"]), dev_f32(sd[f"{p}.norm1.bias"])
            L.wqkv, L.bqkv = dev_bf16(wqkv), dev_f32(bqkv)
            L.wo, L.bo = dev_bf16(sd[f"{p}.self_attn.out_proj.weight"]), dev_f32(sd[f"{p}.self_attn.out_proj.bias"])
            w1, b1 = folded(sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"], sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])
            L.w1, L.b1 = dev_bf16(w1), dev_f32(b1)
            L.w2, L.b2 = dev_bf16(sd[f"{p}.linear2.weight"]), dev_f32(sd[f"{p}.linear2.bias"])
            L.ln2_g, L.ln2_b = dev_f32(ones), dev_f32(zeros)
        W.enc_ln_g, W.enc_ln_b = dev_f32(sd["enc_ln.weight"]), dev_f32(sd["enc_ln.bias"])
        W.ctc_ln_g, W.ctc_ln_b = dev_f32(sd["ctc_head.0.weight"]), dev_f32(sd["ctc_head.0.bias"])
        cw = torch.zeros(self.Cp, D)
        cb = torch.zeros(self.Cp)
        cw[: self.C], cb[: self.C] = sd["ctc_head.2.weight"].float(), sd["ctc_head.2.bias"].float()
        W.ctc_w, W.ctc_b = dev_bf16(cw), dev_f32(cb)

        # ---- decoder
        wmp = sd["mem_proj.weight"].double()
        kvw, kvb = [], []
        for l in range(dec_layers):
            p = f"dec.layers.{l}.multihead_attn"
            wi, bi = sd[f"{p}.in_proj_weight"].double(), sd[f"{p}.in_proj_bias"].float()
            kvw.append((wi[Dd:] @ wmp).float())                  # [(K;V), D_enc]
            kvb.append(bi[Dd:])
        W.crosskv_w, W.crosskv_b = dev_bf16(torch.cat(kvw, 0)), dev_f32(torch.cat(kvb, 0))
        W.dec_emb = dev_f32(sd["dec_emb.weight"])
        pe = sd["dec_pos_enc.pe"][0].float() if self.has_dec_pos else torch.zeros(max_pos, Dd)
        max_pos = pe.shape[0]
        W.dec_pe = dev_f32(pe)
        for l in range(dec_layers):
            p, L = f"dec.layers.{l}", W.dec[l]
            L.wqkv, L.bqkv = dev_bf16(sd[f"{p}.self_attn.in_proj_weight"]), dev_f32(sd[f"{p}.self_attn.in_proj_bias"])
            L.wo, L.bo = dev_bf16(sd[f"{p}.self_attn.out_proj.weight"]), dev_f32(sd[f"{p}.self_attn.out_proj.bias"])
            L.wcq = dev_bf16(sd[f"{p}.multihead_attn.in_proj_weight"][:Dd])
            L.bcq = dev_f32(sd[f"{p}.multihead_attn.in_proj_bias"][:Dd])
            L.wco = dev_bf16(sd[f"{p}.multihead_attn.out_proj.weight"])
            L.bco = dev_f32(sd[f"{p}.multihead_attn.out_proj.bias"])
            L.w1, L.b1 = dev_bf16(sd[f"{p}.linear1.weight"]), dev_f32(sd[f"{p}.linear1.bias"])
            L.w2, L.b2 = dev_bf16(sd[f"{p}.linear2.weight"]), dev_f32(sd[f"{p}.linear2.bias"])
            for i in (1, 2, 3):
                setattr(L, f"ln{i}_g", dev_f32(sd[f"{p}.norm{i}.weight"]))
                setattr(L, f"ln{i}_b", dev_f32(sd[f"{p}.norm{i}.bias"]))
        W.dec_ln_g, W.dec_ln_b = dev_f32(sd["dec_ln.weight"]), dev_f32(sd["dec_ln.bias"])
        hw = torch.zeros(2 * self.Vp, Dd)
        hb = torch.zeros(2 * self.Vp)
        hw[: self.Vd], hb[: self.Vd] = sd["dec_head.weight"].float(), sd["dec_head.bias"].float()
        if self.has_lm:
            hw[self.Vp: self.Vp + self.Vd] = sd["lm_head.weight"].float()
            hb[self.Vp: self.Vp + self.Vd] = sd["lm_head.bias"].float()
        W.heads_w, W.heads_b = dev_bf16(hw), dev_f32(hb)
        self.struct = W

        d = _lib.KiriDims()
        d.img_h, d.enc_dim, d.enc_layers, d.enc_heads, d.enc_ff = cfg.IMG_H, D, enc_layers, cfg.ENC_HEADS, cfg.ENC_FF
        d.dec_dim, d.dec_layers, d.dec_heads, d.dec_ff = Dd, dec_layers, cfg.DEC_HEADS, cfg.DEC_FF
        d.ctc_classes, d.dec_vocab = self.C, self.Vd
        d.max_pos, d.max_t, d.has_dec_pos = max_pos, max_t, int(self.has_dec_pos)
        self.dims = d
        self.enc_layers, self.dec_layers = enc_layers, dec_layers

    def _dev(self, t: torch.Tensor) -> int:
        t = t.contiguous().to(self.device)
        self._keep.append(t)
        return t.data_ptr()
