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
    """Character vocabulary with the reference's three id spaces (contract of kiri_ocr/model.py:83-144: same
    constructor, attribute names and method results; the implementation here is table-driven).

    raw id r in [0, V); CTC id = r + 2 (0 blank, 1 pad); decoder id = r + 3 (0 pad, 1 bos, 2 eos).
    ``vocab.json`` maps token -> stored id; tokens are re-numbered densely in the order of their stored ids and
    ``<unk>`` is appended when the file has none.
    """

    blank_id, pad_id, ctc_offset = 0, 1, 2
    dec_pad, dec_bos, dec_eos, dec_offset = 0, 1, 2, 3

    def __init__(self, vocab_path: str, cfg: CFG):
        with open(vocab_path, "r", encoding="utf-8") as f:
            stored: Dict[str, int] = json.load(f)
        self.unk_token = cfg.UNK_TOKEN
        if self.unk_token not in stored:
            stored[self.unk_token] = 1 + max(stored.values(), default=-1)
        ordered = [tok for tok, _ in sorted(stored.items(), key=lambda item: item[1])]     # stable, like the reference
        self.token_to_id = {tok: raw for raw, tok in enumerate(ordered)}
        self.id_to_token = dict(enumerate(ordered))
        self.unk_id = self.token_to_id[self.unk_token]
        self.vocab_size = len(ordered)
        self.ctc_classes = self.vocab_size + self.ctc_offset
        self.dec_vocab = self.vocab_size + self.dec_offset
        # what every id of each id space contributes to the text: '' for specials and <unk>
        emitted = ["" if tok == self.unk_token else tok for tok in ordered]
        self.ctc_text = [""] * self.ctc_offset + emitted
        self.dec_text = [""] * self.dec_offset + emitted

    def decode_ctc(self, ids: List[int]) -> str:
        """Frame ids -> text: runs of equal ids count once, then blank / pad / <unk> / out-of-range ids vanish."""
        from itertools import groupby
        table, n = self.ctc_text, self.ctc_classes
        return "".join(table[i] for i, _ in groupby(ids) if 0 <= i < n)

    def decode_collapsed_ctc(self, ids: List[int]) -> str:
        """Text for ids the device already collapsed (repeats removed, ids >= 2 kept) — the tail of ``decode_ctc``."""
        table, n = self.ctc_text, self.ctc_classes
        return "".join(table[i] for i in ids if 0 <= i < n)

    def decode_dec(self, ids: List[int]) -> str:
        table, n = self.dec_text, self.dec_vocab
        return "".join(table[i] for i in ids if 0 <= i < n)

    # ---- batched decoding (host side of the batched engine: one vectorised pass instead of a Python loop per id)
    def _codes(self, space: str):
        """uint32 code point emitted by every id of an id space (0 = nothing), or None when some token is not a
        single encodable character (then the per-line path is used)."""
        cache = self.__dict__.setdefault("_code_cache", {})
        if space not in cache:
            import numpy as np
            table = self.ctc_text if space == "ctc" else self.dec_text
            ok = all(len(t) <= 1 and not (t and 0xD800 <= ord(t) <= 0xDFFF) and t != "\x00" for t in table)
            # one extra 0 entry at the end: every id outside the table is clipped onto it
            cache[space] = np.array([ord(t) if t else 0 for t in table] + [0], np.uint32) if ok else None
        return cache[space]

    def decode_batch(self, ids_flat, lengths, space: str = "ctc") -> List[str]:
        """Texts of many lines at once: ``ids_flat`` holds the lines' ids back to back (already collapsed CTC ids, or
        decoder ids already cut before EOS), ``lengths[i]`` ids belong to line i.  Same result as
        ``decode_collapsed_ctc`` / ``decode_dec`` per line."""
        import numpy as np
        ids_flat = np.asarray(ids_flat, np.int64)
        lengths = np.asarray(lengths, np.int64)
        codes_tab = self._codes(space)
        if codes_tab is None:
            one = self.decode_collapsed_ctc if space == "ctc" else self.decode_dec
            ends = np.cumsum(lengths)
            return [one(ids_flat[e - n:e].tolist()) for n, e in zip(lengths.tolist(), ends.tolist())]
        # negative ids become huge as unsigned, so ONE minimum sends every id outside the table to its empty last entry
        codes = np.take(codes_tab, np.minimum(ids_flat.view(np.uint64), np.uint64(len(codes_tab) - 1)).view(np.int64))
        keep = codes != 0
        line = np.repeat(np.arange(len(lengths)), lengths)
        kept = np.bincount(line[keep], minlength=len(lengths))
        big = codes[keep].astype("<u4").tobytes().decode("utf-32-le")
        ends = np.cumsum(kept)
        return [big[e - n:e] for n, e in zip(kept.tolist(), ends.tolist())]

    def dec_to_ctc_id(self, dec_id: int) -> int:
        """Decoder id -> CTC id: specials map to blank, anything outside the vocabulary to <unk>'s CTC id."""
        if dec_id < self.dec_offset:
            return self.blank_id if dec_id >= 0 else self.unk_id + self.ctc_offset
        if dec_id < self.dec_vocab:
            return dec_id - self.dec_offset + self.ctc_offset
        return self.unk_id + self.ctc_offset
