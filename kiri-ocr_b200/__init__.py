"""kiri-ocr_b200 — B200-native batched text-line recognition behind kiri-ocr's OCR API.

Scope: crop resize/normalise -> conv stem -> transformer encoder -> CTC greedy
("fast") or greedy KV-cached attention decoder ("accurate"), as hand-written sm_100a CUDA
kernels behind a C ABI (``include/kiri_b200.h``), driven by Python/PyTorch host code that keeps
the reference's ``OCR`` class API, detector plug-in boundary and checkpoint layout.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy, like the reference package (kiri_ocr/__init__.py:15-35)
    if name == "OCR":
        from .core import OCR
        return OCR
    if name in ("CFG", "CharTokenizer"):
        from . import config
        return getattr(config, name)
    if name == "BatchedRecognizer":
        from .engine import BatchedRecognizer
        return BatchedRecognizer
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


__all__ = ["OCR", "CFG", "CharTokenizer", "BatchedRecognizer"]
