"""Oracle: line-crop preprocessing in exact integer arithmetic (numpy).  Test infrastructure.

Restates
  * ``OCR._preprocess_region``                — kiri_ocr/core.py:489-528
  * ``ResizeKeepRatioPadNoCrop.__call__``     — kiri_ocr/model.py:316-331
  * ``preprocess_pil``                        — kiri_ocr/model.py:334-339
  * Pillow ``Image.resize(BILINEAR)`` for mode "L" (third-party, Pillow>=9.0.0 per
    pyproject.toml:17; 12.2.0 installed): separable triangle filter with antialias support
    ``max(1, scale)``, horizontal pass then vertical pass, weights normalised in float64 and
    quantised to ``int(0.5 + w * 2**22)``, each pass accumulated in integers, rounded with
    ``+ 2**21 >> 22`` and clipped to uint8 (SURVEY.md §8a pseudo-code).
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


@functools.lru_cache(maxsize=4096)
def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per output index: first tap ``xmin``, tap count ``n`` and integer weights ``k[out, ksize]``
    (zero padded), following Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc``.
    Vectorised over the output index; the weight sum runs tap by tap in the C loop's order
    (trailing zero taps do not change a float64 sum), so the result is bit-identical."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                      # bilinear filter support is 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5), 0.0)            # C (int) truncation
    xmax = np.minimum(np.trunc(center + support + 0.5), float(in_size))
    n = (xmax - xmin).astype(np.int64)
    w = np.zeros((out_size, ksize), np.float64)
    ww = np.zeros(out_size, np.float64)
    for k in range(ksize):
        arg = np.abs((k + xmin - center + 0.5) * ss)
        wk = np.where((k < n) & (arg < 1.0), 1.0 - arg, 0.0)
        w[:, k] = wk
        ww = ww + wk
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    q = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
    kk = np.trunc(q).astype(np.int64)
    kk[np.arange(ksize)[None, :] >= n[:, None]] = 0
    return xmin.astype(np.int32), n.astype(np.int32), kk


def _pass_last_axis(a: np.ndarray, out_size: int) -> np.ndarray:
    """One resample pass along the last axis of a uint8 array."""
    in_size = a.shape[-1]
    xmins, counts, kk = resample_coeffs(in_size, out_size)
    ksize = kk.shape[1]
    idx = xmins[:, None] + np.arange(ksize)[None, :]
    idx = np.minimum(idx, in_size - 1)                # padded taps carry weight 0
    g = a[..., idx].astype(np.int64)                  # [..., out, ksize]
    acc = (g * kk).sum(-1) + (1 << (PRECISION_BITS - 1))
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_bilinear_resize(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR)`` for a uint8 2-D array."""
    h, w = img.shape
    t = img
    if out_w != w:
        t = _pass_last_axis(t, out_w)                                   # horizontal first
    if out_h != h:
        t = np.ascontiguousarray(_pass_last_axis(np.ascontiguousarray(t.T), out_h).T)
    return t


def target_width(iw: int, ih: int, img_h: int = 48) -> int:
    """model.py:321-322 — Python ``round`` is round-half-to-even."""
    scale = img_h / float(ih)
    return max(1, int(round(iw * scale)))


def resize_keep_ratio_pad(img: np.ndarray, img_h: int = 48, img_w: int = 640) -> np.ndarray:
    """model.py:316-331: resize to height ``img_h``, crop to ``img_w`` or left-paste on gray 128."""
    ih, iw = img.shape
    nw = target_width(iw, ih, img_h)
    r = pil_bilinear_resize(img, nw, img_h)
    if nw >= img_w:
        return np.ascontiguousarray(r[:, :img_w])
    out = np.full((img_h, img_w), 128, np.uint8)
    out[:, :nw] = r
    return out


def crop_region(page: np.ndarray, box, extra_padding: int = 5) -> Optional[np.ndarray]:
    """core.py:506-525: clamp-pad the box, slice, invert when the mean is below 127."""
    img_h, img_w = page.shape[:2]
    x, y, w, h = (int(v) for v in box)
    x1 = max(0, x - extra_padding)
    y1 = max(0, y - extra_padding)
    x2 = min(img_w, x + w + extra_padding)
    y2 = min(img_h, y + h + extra_padding)
    roi = page[y1:y2, x1:x2]
    if roi.size == 0:
        return None
    if int(roi.astype(np.int64).sum()) < 127 * roi.size:    # == (np.mean(roi) < 127)
        roi = 255 - roi
    return roi


def normalise(plane_u8: np.ndarray) -> np.ndarray:
    """model.py:337-338, in the reference's op order: ``/255`` then ``(x-0.5)/0.5`` in fp32."""
    x = plane_u8.astype(np.float32) / np.float32(255.0)
    return (x - np.float32(0.5)) / np.float32(0.5)


def preprocess_region(page: np.ndarray, box, img_h: int = 48, img_w: int = 640,
                      extra_padding: int = 5) -> Optional[np.ndarray]:
    """uint8 ``[img_h, img_w]`` plane for one box, or None for an empty crop."""
    roi = crop_region(page, box, extra_padding)
    if roi is None:
        return None
    return resize_keep_ratio_pad(roi, img_h, img_w)


def preprocess_crop(crop: np.ndarray, img_h: int = 48, img_w: int = 640) -> np.ndarray:
    """A stand-alone crop (already cut out of its page): invert-if-dark, resize, crop/pad."""
    return resize_keep_ratio_pad(crop_region(crop, (0, 0, crop.shape[1], crop.shape[0]), 0), img_h, img_w)
