"""Oracle: decode algorithms (CTC greedy, greedy attention decoder).  Test infrastructure.

Restates
  * ``compute_ctc_confidence``           — kiri_ocr/model.py:343-373
  * ``CharTokenizer.decode_ctc``         — kiri_ocr/model.py:109-124
  * ``beam_decode_one_batched`` at BEAM=1 ("accurate", core.py:560-568) — kiri_ocr/model.py:390-600
  * ``compute_sequence_confidence``      — kiri_ocr/model.py:376-386
  * ``greedy_ctc_decode_streaming``      — kiri_ocr/model.py:689-775
  * ``greedy_decode_streaming``          — kiri_ocr/model.py:779-946
  * ``OCR.recognize_region`` dispatch    — kiri_ocr/core.py:530-575
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import model as M


# --------------------------------------------------------------------------- CTC greedy
def ctc_greedy(logits: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float, int]:
    """[T,C] logits -> (frame argmax ids [T], collapsed ids (repeats removed, ids>=2), mean
    max-softmax confidence, length estimate).  model.py:355-371 and 109-119."""
    x = torch.as_tensor(logits, dtype=torch.float32)
    probs = F.softmax(x, dim=-1)
    best = x.argmax(dim=-1).numpy()
    conf = float(probs.max(dim=-1).values.mean().item())
    collapsed: List[int] = []
    prev = None
    length = 0
    for idx in best.tolist():
        if idx != prev and idx >= 2:
            collapsed.append(idx)
            length += 1
        prev = idx
    return best, np.asarray(collapsed, np.int32), conf, length


def max_steps_for(cfg, target_len: int, mem_len: int) -> int:
    """model.py:416-425."""
    if target_len and target_len > 0:
        return min(cfg.MAX_DEC_LEN, int(target_len * cfg.DEC_MAX_LEN_RATIO) + cfg.DEC_MAX_LEN_PAD)
    return min(cfg.MAX_DEC_LEN, int(mem_len * cfg.MEM_MAX_LEN_RATIO) + cfg.DEC_MAX_LEN_PAD)


def apply_penalties(logp: torch.Tensor, seq: List[int], cfg, unk_dec_id: int,
                    target_len: Optional[int]) -> None:
    """In-place EOS bias, the four cumulative repeat penalties and the UNK penalty for one
    hypothesis whose ids so far (BOS included) are ``seq`` — model.py:490-534."""
    cur_len = len(seq) - 1
    if target_len and target_len > 0:
        min_len = min(cfg.EOS_BIAS_UNTIL_LEN, max(1, int(target_len * 0.5)))
        if cur_len < min_len:
            logp[2] -= cfg.EOS_LOGP_BIAS
        elif cur_len >= target_len:
            logp[2] += cfg.EOS_LOGP_BOOST
    elif cur_len < cfg.EOS_BIAS_UNTIL_LEN:
        logp[2] -= cfg.EOS_LOGP_BIAS
    n = len(seq)
    if n >= 4 and seq[-1] == seq[-2] == seq[-3]:
        logp[seq[-1]] -= cfg.REPEAT_LAST_PENALTY
    if n >= 4 and (seq[-2], seq[-1]) == (seq[-4], seq[-3]):
        logp[seq[-1]] -= cfg.REPEAT_BIGRAM_PENALTY
        logp[seq[-2]] -= cfg.REPEAT_BIGRAM_PENALTY
    if n >= 3 and seq[-1] == seq[-3]:
        if n >= 4 and seq[-2] == seq[-4]:
            logp[seq[-1]] -= cfg.REPEAT_BIGRAM_PENALTY
    if n >= 6 and (seq[-3], seq[-2], seq[-1]) == (seq[-6], seq[-5], seq[-4]):
        logp[seq[-1]] -= cfg.REPEAT_TRIGRAM_PENALTY
        logp[seq[-2]] -= cfg.REPEAT_TRIGRAM_PENALTY
        logp[seq[-3]] -= cfg.REPEAT_TRIGRAM_PENALTY
    logp[unk_dec_id] -= cfg.UNK_LOGP_PENALTY


def fused_logp(dec_logits: torch.Tensor, lm_logits: Optional[torch.Tensor], cfg) -> torch.Tensor:
    """model.py:480-485."""
    logp = F.log_softmax(dec_logits, dim=-1)
    if cfg.USE_LM and cfg.USE_LM_FUSION_EVAL and lm_logits is not None:
        logp = logp + cfg.LM_FUSION_ALPHA * F.log_softmax(lm_logits, dim=-1)
    return logp


@torch.inference_mode()
def greedy_decode(sd, memp_1: torch.Tensor, cfg, unk_dec_id: int, target_len: int,
                  heads: int = 8, forced: Optional[List[int]] = None,
                  return_logp: bool = False):
    """Greedy attention decode of ONE line (``beam_decode_one_batched`` with BEAM=1).

    Returns (ids chosen incl. a final EOS if emitted, their penalised log-probs).  With
    ``forced`` the given token sequence is fed instead of the arg-max (teacher forcing) and
    ``return_logp`` also returns the per-step penalised log-prob rows."""
    st = M.DecoderState(sd, memp_1, heads)
    steps = max_steps_for(cfg, target_len, memp_1.shape[1])
    seq = [1]
    out_ids: List[int] = []
    out_lp: List[float] = []
    rows = []
    for step in range(steps):
        dec, lm = M.decoder_step(st, torch.tensor([seq[-1]]))
        logp = fused_logp(dec, lm, cfg)[0].clone()
        apply_penalties(logp, seq, cfg, unk_dec_id, target_len)
        if return_logp:
            rows.append(logp.clone())
        if forced is not None:
            if step >= len(forced):
                break
            tid = int(forced[step])
        else:
            tid = int(torch.topk(logp, 1).indices[0])
        out_ids.append(tid)
        out_lp.append(float(logp[tid]))
        seq.append(tid)
        if tid == 2:
            break
    if return_logp:
        return out_ids, out_lp, (torch.stack(rows) if rows else torch.zeros(0))
    return out_ids, out_lp


def sequence_confidence(log_probs: List[float]) -> float:
    if not log_probs:
        return 0.0
    return min(1.0, max(0.0, math.exp(sum(log_probs) / len(log_probs))))


# --------------------------------------------------------------------------- line level
@torch.inference_mode()
def recognize_plane(sd, tok, cfg, plane_u8: np.ndarray, method: str = "ctc",
                    heads: int = 8) -> Tuple[str, float, Dict]:
    """``OCR.recognize_region`` on one preprocessed uint8 plane (core.py:530-575): returns
    (text, confidence, details) for method in {"ctc", "decoder"}."""
    from .preprocess import normalise

    x = torch.from_numpy(normalise(plane_u8))[None, None]
    mem = M.encode(sd, x, heads)
    logits = M.ctc_logits(sd, mem)[0]
    best, collapsed, ctc_conf, length = ctc_greedy(logits.numpy())
    info = {"ctc_ids": collapsed, "ctc_conf": ctc_conf, "len_est": length}
    if method == "ctc":
        return tok.decode_ctc(best.tolist()), ctc_conf, info
    memp = M.mem_proj(sd, mem)
    ids, lps = greedy_decode(sd, memp, cfg, tok.unk_id + tok.dec_offset, length, heads)
    text_ids = []
    for t in ids:
        if t == tok.dec_eos:
            break
        text_ids.append(t)
    conf = 0.6 * sequence_confidence(lps) + 0.4 * ctc_conf            # model.py:594-598
    info.update(dec_ids=np.asarray(ids, np.int32), dec_logps=np.asarray(lps, np.float64))
    return tok.decode_dec(text_ids), conf, info


# --------------------------------------------------------------------------- streaming forms
def ctc_stream_chunks(logits: np.ndarray, tok) -> Iterator[Dict]:
    """``greedy_ctc_decode_streaming`` (model.py:719-775) from [T,C] logits."""
    x = torch.as_tensor(logits, dtype=torch.float32)
    probs = F.softmax(x, dim=-1)
    best = x.argmax(dim=-1).tolist()
    maxp = probs.max(dim=-1).values
    text, prev, step = "", None, 0
    for t, idx in enumerate(best):
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
                yield {"token": ch, "token_id": idx, "text": text, "confidence": float(maxp[t]),
                       "step": step, "finished": False}
    yield {"token": "", "token_id": -1, "text": text, "confidence": float(maxp.mean()),
           "step": step, "finished": True}


@torch.inference_mode()
def greedy_stream_chunks(sd, memp_1, tok, cfg, target_len: int, heads: int = 8) -> Iterator[Dict]:
    """``greedy_decode_streaming`` (model.py:779-946): the token is the arg-max of the RAW
    ``dec_head`` softmax (915-917); fusion and penalties only shape the recorded log-prob."""
    st = M.DecoderState(sd, memp_1, heads)
    steps = max_steps_for(cfg, target_len, memp_1.shape[1])
    seq, text = [1], ""
    for step in range(steps):
        dec, lm = M.decoder_step(st, torch.tensor([seq[-1]]))
        probs = F.softmax(dec, dim=-1)[0]
        best_prob, best_id = probs.max(dim=0)
        best_id = int(best_id)
        finished = best_id == tok.dec_eos
        ch = ""
        if not finished and best_id not in (0, 1, 2):
            raw = best_id - tok.dec_offset
            if 0 <= raw < tok.vocab_size:
                ch = tok.id_to_token.get(raw, "")
                if ch != tok.unk_token:
                    text += ch
        seq.append(best_id)
        yield {"token": ch, "token_id": best_id, "text": text, "confidence": float(best_prob),
               "step": step + 1, "finished": finished}
        if finished:
            break
