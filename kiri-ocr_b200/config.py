"""Recognizer configuration and tokenizer — host-side mirror of the reference contract.

Mirrors ``kiri_ocr/model.py:24-69`` (``CFG``: architecture + decode hyper-parameters) and
``kiri_ocr/model.py:83-144`` (``CharTokenizer``: id spaces).  Field names, defaults and id
conventions are the reference's, because checkpoints (``*_meta.json``) and user code
(``ocr.cfg.BEAM = 5``) address them by name.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Dict, List


@dataclass
class CFG:
    # architecture (model.py:27-44)
    IMG_H: int = 48
    IMG_W: int = 640
    MAX_DEC_LEN: int = 512
    UNK_TOKEN: str = "<unk>"
    COLLAPSE_WHITESPACE: bool = True
    UNICODE_NFC: bool = True

    ENC_DIM: int = 256
    ENC_LAYERS: int = 4
    ENC_HEADS: int = 8
    ENC_FF: int = 1024
    DROPOUT: float = 0.15

    USE_DECODER: bool = True
    DEC_DIM: int = 256
    DEC_LAYERS: int = 3
    DEC_HEADS: int = 8
    DEC_FF: int = 1024

    USE_CTC: bool = True
    USE_LM: bool = True
    USE_LM_FUSION_EVAL: bool = True
    LM_FUSION_ALPHA: float = 0.35
    USE_FP16: bool = True
    USE_AUTOCAST: bool = True

    # inference parameters (model.py:53-69)
    CTC_FUSION_ALPHA: float = 0.5
    BEAM: int = 3
    BEAM_LENP: float = 0.8

    EOS_LOGP_BIAS: float = 0.0
    EOS_LOGP_BOOST: float = 0.0
    EOS_BIAS_UNTIL_LEN: int = 2

    REPEAT_LAST_PENALTY: float = 3
    REPEAT_BIGRAM_PENALTY: float = 2.5
    REPEAT_TRIGRAM_PENALTY: float = 2.0
    UNK_LOGP_PENALTY: float = 10

    DEC_MAX_LEN_RATIO: float = 1.3
    DEC_MAX_LEN_PAD: int = 10
    MEM_MAX_LEN_RATIO: float = 1


# keys written by training.py:1013-1038 and read back by core.py:425-450
META_CONFIG_KEYS = (
    "IMG_H", "IMG_W", "ENC_DIM", "ENC_LAYERS", "ENC_HEADS", "ENC_FF",
    "DEC_DIM", "DEC_LAYERS", "DEC_HEADS", "DEC_FF", "DROPOUT", "USE_CTC", "USE_FP16",
)


class CharTokenizer:
    """Character vocabulary with the reference's three id spaces (model.py:83-144).

    raw id r in [0, V); CTC id = r + 2 (0 blank, 1 pad); decoder id = r + 3 (0 pad, 1 bos, 2 eos).
    """

    def __init__(self, vocab_path: str, cfg: CFG):
        with open(vocab_path, "r", encoding="utf-8") as f:
            vocab_raw: Dict[str, int] = json.load(f)
        if cfg.UNK_TOKEN not in vocab_raw:
            vocab_raw[cfg.UNK_TOKEN] = max(vocab_raw.values(), default=-1) + 1
        # ids are re-densified in order of their stored value (model.py:91-93)
        items = sorted(vocab_raw.items(), key=lambda kv: kv[1])
        self.token_to_id = {tok: i for i, (tok, _) in enumerate(items)}
        self.id_to_token = {i: tok for i, (tok, _) in enumerate(items)}

        self.unk_token = cfg.UNK_TOKEN
        self.unk_id = self.token_to_id[cfg.UNK_TOKEN]
        self.blank_id = 0
        self.pad_id = 1
        self.ctc_offset = 2
        self.vocab_size = len(self.token_to_id)
        self.ctc_classes = self.vocab_size + self.ctc_offset

        self.dec_pad = 0
        self.dec_bos = 1
        self.dec_eos = 2
        self.dec_offset = 3
        self.dec_vocab = self.vocab_size + self.dec_offset

    def decode_ctc(self, ids: List[int]) -> str:
        """Collapse repeats, then drop blank/pad and <unk> (model.py:109-124)."""
        chars = []
        prev_id = None
        for idx in ids:
            if idx == prev_id:
                continue
            prev_id = idx
            if idx < self.ctc_offset:
                continue
            raw_id = idx - self.ctc_offset
            if 0 <= raw_id < self.vocab_size:
                ch = self.id_to_token.get(raw_id, "")
                if ch != self.unk_token:
                    chars.append(ch)
        return "".join(chars)

    def decode_collapsed_ctc(self, ids: List[int]) -> str:
        """Text for ids that the device already collapsed (repeats removed, ids >= 2 kept).

        Equivalent to ``decode_ctc`` on the un-collapsed frame ids: the device kernel applies
        the ``idx == prev`` and ``idx < 2`` rules, this applies the range and <unk> rules.
        """
        chars = []
        for idx in ids:
            raw_id = idx - self.ctc_offset
            if 0 <= raw_id < self.vocab_size:
                ch = self.id_to_token.get(raw_id, "")
                if ch != self.unk_token:
                    chars.append(ch)
        return "".join(chars)

    def decode_dec(self, ids: List[int]) -> str:
        out = []
        for x in ids:
            if x in (self.dec_pad, self.dec_bos, self.dec_eos):
                continue
            y = x - self.dec_offset
            if 0 <= y < self.vocab_size:
                t = self.id_to_token.get(y, self.unk_token)
                out.append("" if t == self.unk_token else t)
        return "".join(out)

    def dec_to_ctc_id(self, dec_id: int) -> int:
        if dec_id in (self.dec_pad, self.dec_bos, self.dec_eos):
            return self.blank_id
        raw_id = dec_id - self.dec_offset
        if 0 <= raw_id < self.vocab_size:
            return raw_id + self.ctc_offset
        return self.unk_id + self.ctc_offset
