"""Oracle: decode algorithms (CTC greedy, greedy attention decoder).  Test infrastructure.

Restates
  * ``compute_ctc_confidence``           — kiri_ocr/model.py:343-373
  * ``CharTokenizer.decode_ctc``         — kiri_ocr/model.py:109-124
  * ``beam_decode_one_batched`` at BEAM=1 ("accurate", core.py:560-568) — kiri_ocr/model.py:390-600
  * ``compute_sequence_confidence``      — kiri_ocr/model.py:376-386
  * ``greedy_ctc_decode_streaming``      — kiri_ocr/model.py:689-775
  * ``greedy_decode_streaming``          — kiri_ocr/model.py:779-946
  * ``OCR.recognize_region`` dispatch    — kiri_ocr/core.py:530-575
  * ``beam_decode_one_batched`` at BEAM>1 ("beam") — kiri_ocr/model.py:390-600
  * ``compute_ctc_alignment_score``      — kiri_ocr/model.py:603-668
  * ``beam_decode_streaming``            — kiri_ocr/model.py:949-1152
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


# --------------------------------------------------------------------------- beam search
def ctc_alignment_score(ctc_logits: torch.Tensor, dec_seq: List[int], tok) -> float:
    """model.py:603-668.  The reference evaluates one ``torch.logsumexp`` per (t, s); here the 1/2/3
    candidate rule is vectorised over s with -inf for the absent candidates, which gives the same
    fp32 values (a zero term does not change the 3-way sum)."""
    if ctc_logits.dim() == 3:
        ctc_logits = ctc_logits.squeeze(0)
    log_probs = F.log_softmax(ctc_logits.float(), dim=-1)
    labels = []
    for x in dec_seq[1:]:
        if x == tok.dec_eos:
            break
        if x in (tok.dec_pad, tok.dec_bos):
            continue
        labels.append(tok.dec_to_ctc_id(x))
    if not labels:
        return log_probs[:, tok.blank_id].sum().item() / max(1, log_probs.size(0))
    T, blank = log_probs.size(0), tok.blank_id
    ext = [blank]
    for lid in labels:
        ext += [lid, blank]
    S = len(ext)
    ext_t = torch.tensor(ext)
    ninf = float("-inf")
    skip_ok = torch.zeros(S, dtype=torch.bool)
    for s_ in range(2, S):
        skip_ok[s_] = ext[s_] != blank and ext[s_] != ext[s_ - 2]
    alpha = log_probs.new_full((S,), ninf)
    alpha[0] = log_probs[0, blank]
    if S > 1:
        alpha[1] = log_probs[0, ext[1]]
    for t in range(1, T):
        a1 = torch.cat([alpha.new_full((1,), ninf), alpha[:-1]])
        a2 = torch.cat([alpha.new_full((2,), ninf), alpha[:-2]])
        a2 = torch.where(skip_ok, a2, alpha.new_full((S,), ninf))
        stacked = torch.stack([alpha, a1, a2])
        m = stacked.max(dim=0).values
        lse = torch.where(torch.isinf(m), m, m + torch.log(torch.exp(stacked - m).sum(dim=0)))
        alpha = lse + log_probs[t, ext_t]
    if S == 1:
        total = alpha[0]
    else:
        total = torch.logsumexp(torch.stack([alpha[S - 1], alpha[S - 2]]), dim=0)
    return total.item() / max(1, len(labels))


def _fork_state(st: "M.DecoderState", parents: List[int]) -> "M.DecoderState":
    """KV caches of the hypotheses that descend from ``parents`` (rows of the previous step)."""
    import copy
    idx = torch.tensor(parents, dtype=torch.long)
    new = copy.copy(st)
    new.self_k = [None if k is None else k.index_select(0, idx) for k in st.self_k]
    new.self_v = [None if v is None else v.index_select(0, idx) for v in st.self_v]
    new.cross_k = [c[:1].expand(len(parents), -1, -1).contiguous() for c in st.cross_k]
    new.cross_v = [c[:1].expand(len(parents), -1, -1).contiguous() for c in st.cross_v]
    return new


@torch.inference_mode()
def beam_decode(sd, memp_1: torch.Tensor, ctc_logits_1: torch.Tensor, tok, cfg, heads: int = 8):
    """``beam_decode_one_batched`` (model.py:390-600) with a KV cache instead of the full-prefix
    re-run.  Returns (text, confidence, info) where info carries every surviving hypothesis."""
    _, collapsed, ctc_conf, target_len = ctc_greedy(ctc_logits_1.reshape(-1, ctc_logits_1.shape[-1]).numpy())
    max_steps = max_steps_for(cfg, target_len, memp_1.shape[1])
    unk = tok.unk_id + tok.dec_offset
    beams = [(0.0, [tok.dec_bos], [], False)]
    st = M.DecoderState(sd, memp_1, heads)
    alive_rows: List[int] = [0]                       # cache row of every alive hypothesis
    for step in range(max_steps):
        if all(b[3] for b in beams):
            break
        alive = [b for b in beams if not b[3]]
        done = [b for b in beams if b[3]]
        if not alive:
            beams = done
            break
        st = _fork_state(st, alive_rows)
        dec, lm = M.decoder_step(st, torch.tensor([b[1][-1] for b in alive]))
        logp = fused_logp(dec, lm, cfg).clone()
        for i, (_, seq, _, _) in enumerate(alive):
            apply_penalties(logp[i], seq, cfg, unk, target_len)
        topv, topi = torch.topk(logp, k=cfg.BEAM, dim=-1)
        new_beams = [(b, -1) for b in done]
        for bi, (base, seq, lps, _) in enumerate(alive):
            for v, tid in zip(topv[bi].tolist(), topi[bi].tolist()):
                new_beams.append(((base + float(v), seq + [int(tid)], lps + [float(v)], int(tid) == tok.dec_eos), bi))

        def normed(entry):
            score, seq, _, _ = entry[0]
            L = max(1, len(seq) - 1)
            return score / (((5 + L) ** cfg.BEAM_LENP) / ((5 + 1) ** cfg.BEAM_LENP))

        new_beams.sort(key=normed, reverse=True)
        kept = new_beams[: cfg.BEAM]
        beams = [b for b, _ in kept]
        alive_rows = [row for b, row in kept if not b[3]]

    def final(entry):
        score, seq, lps, _ = entry
        length = max(1, len(seq) - 1)
        dec_score = score / (length ** cfg.BEAM_LENP if length > 0 else 1.0)
        conf = sequence_confidence(lps)
        if cfg.CTC_FUSION_ALPHA > 0:
            return dec_score + cfg.CTC_FUSION_ALPHA * ctc_alignment_score(ctc_logits_1, seq, tok), conf
        return dec_score, conf

    scored = [(final(b), b) for b in beams]
    scored.sort(key=lambda x: x[0][0], reverse=True)
    (_, best_conf), (_, best_seq, _, _) = scored[0]
    ids = []
    for x in best_seq[1:]:
        if x == tok.dec_eos:
            break
        ids.append(x)
    info = {"beams": beams, "scored": [(sc, b[1]) for (sc, _), b in scored], "len_est": target_len, "ctc_conf": ctc_conf}
    return tok.decode_dec(ids), 0.6 * best_conf + 0.4 * ctc_conf, info


@torch.inference_mode()
def beam_stream_chunks(sd, memp_1: torch.Tensor, ctc_logits_1: torch.Tensor, tok, cfg, heads: int = 8) -> Iterator[Dict]:
    """``beam_decode_streaming`` (model.py:949-1152) with a KV cache: same expansion and penalties as ``beam_decode``
    but hypotheses are pruned by ``score / L**BEAM_LENP`` (1112-1115), the best partial hypothesis is yielded after
    every step and decoding stops as soon as the BEST hypothesis has ended (1148-1150); no CTC rescoring."""
    _, _, _, target_len = ctc_greedy(ctc_logits_1.reshape(-1, ctc_logits_1.shape[-1]).numpy())
    max_steps = max_steps_for(cfg, target_len, memp_1.shape[1])
    unk = tok.unk_id + tok.dec_offset
    beams = [(0.0, [tok.dec_bos], [], False)]
    st = M.DecoderState(sd, memp_1, heads)
    alive_rows: List[int] = [0]
    prev = ""
    for step in range(max_steps):
        if all(b[3] for b in beams):
            break
        alive = [b for b in beams if not b[3]]
        done = [b for b in beams if b[3]]
        if not alive:
            break
        st = _fork_state(st, alive_rows)
        dec, lm = M.decoder_step(st, torch.tensor([b[1][-1] for b in alive]))
        logp = fused_logp(dec, lm, cfg).clone()
        for i, (_, seq, _, _) in enumerate(alive):
            apply_penalties(logp[i], seq, cfg, unk, target_len)
        topv, topi = torch.topk(logp, k=cfg.BEAM, dim=-1)
        new_beams = [(b, -1) for b in done]
        for bi, (base, seq, lps, _) in enumerate(alive):
            for v, tid in zip(topv[bi].tolist(), topi[bi].tolist()):
                new_beams.append(((base + float(v), seq + [int(tid)], lps + [float(v)], int(tid) == tok.dec_eos), bi))
        new_beams.sort(key=lambda e: e[0][0] / (max(1, len(e[0][1]) - 1) ** cfg.BEAM_LENP), reverse=True)
        kept = new_beams[: cfg.BEAM]
        beams = [b for b, _ in kept]
        alive_rows = [row for b, row in kept if not b[3]]
        _, best_seq, best_lps, best_fin = beams[0]
        ids = []
        for x in best_seq[1:]:
            if x == tok.dec_eos:
                break
            ids.append(x)
        cur = tok.decode_dec(ids)
        yield {"token": cur[len(prev):] if len(cur) > len(prev) else "", "text": cur,
               "confidence": sequence_confidence(best_lps) if best_lps else 0.0, "step": step + 1, "finished": best_fin,
               "best_ids": list(best_seq[1:])}
        prev = cur
        if best_fin:
            break


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
    if method == "beam":
        text, conf, binfo = beam_decode(sd, memp, logits, tok, cfg, heads)
        info.update(binfo)
        return text, conf, info
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
