"""Batched validation forward on the B200 engine.

Mirror of the validation block of the reference's training loop (kiri_ocr/training.py:865-949): for every batch of the
validation loader the CTC greedy text of EVERY sample is compared with its label (exact match after ``strip``), and the
greedy attention decoder (``cfg.BEAM = 1``) is sampled on the FIRST sample of every 10th batch.  The reference encodes
the batch once (``model.encode(imgs)``, training.py:891-895) and then walks it sample by sample on the host; here the
whole batch goes through the engine in one call (preprocess is the identity on already normalised planes, no
inversion), and the decoder samples of the epoch share the batch they belong to.

A loader yields dicts with ``"images"``: float tensor ``[B, 1, IMG_H, W]`` in [-1, 1] as ``preprocess_pil`` produces
them (kiri_ocr/model.py:334-339) and ``"texts"``: list of ``B`` strings.  Values are mapped back to the uint8 plane they
came from; images that are not uint8-representable (augmented tensors) are rounded to the nearest level and counted
in ``inexact_images``.  There is no backward pass: training stays with the reference.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib


def _planes_u8(images: torch.Tensor):
    t = images.detach().float().cpu()
    if t.dim() != 4 or t.shape[1] != 1:
        raise ValueError("images must be [B, 1, IMG_H, W]")
    v = (t[:, 0] * 0.5 + 0.5) * 255.0
    q = torch.round(v).clamp(0, 255)
    inexact = int(((v - q).abs().amax(dim=(1, 2)) > 1e-3).sum())
    return q.to(torch.uint8).numpy(), inexact


def validate_recognizer(engine, val_loader: Iterable[Dict], max_val_samples: Optional[int] = None,
                        decoder_every: int = 10) -> Dict[str, float]:
    """Returns the reference's validation numbers: ``val_acc`` (CTC exact-match %, training.py:931),
    ``val_dec_acc`` (sampled decoder %, training.py:933-934) and the raw counts."""
    total = ctc_correct = dec_correct = inexact = 0
    batch_idx = -1
    for batch_idx, batch in enumerate(val_loader):
        if max_val_samples and total >= max_val_samples:
            break
        planes, bad = _planes_u8(batch["images"])
        texts = batch["texts"]
        inexact += bad
        B, H, W = planes.shape
        if max_val_samples:
            B = min(B, max_val_samples - total)
        ent = np.array([(i * H * W, W, W, H, _lib.CROP_NO_INVERT) for i in range(B)], np.int64)
        flat = [planes.reshape(-1)]
        res = engine.recognize_packed(flat, ent, "ctc")
        for i in range(B):
            if res[i].text.strip() == texts[i].strip():
                ctc_correct += 1
            total += 1
        if batch_idx % decoder_every == 0:
            dec = engine.recognize_packed(flat, ent[:1], "decoder")[0]
            if dec.text.strip() == texts[0].strip():
                dec_correct += 1
    sampled_batches = (batch_idx + 1) // decoder_every + 1
    return {"val_acc": ctc_correct / max(1, total) * 100.0, "val_dec_acc": dec_correct / max(1, sampled_batches) * 100.0,
            "val_total": total, "val_ctc_correct": ctc_correct, "val_dec_correct": dec_correct,
            "sampled_batches": sampled_batches, "inexact_images": inexact}
