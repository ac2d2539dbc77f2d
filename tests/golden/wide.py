"""The WIDE-MARGIN fixtures: checkpoints whose output heads are FITTED so that every CTC frame and every
decoder step of the fixture lines has a top-1 margin far above the bf16 tolerance.  Test infrastructure.

Why fitted and not "head gain": scaling ``ctc_head.2`` / ``dec_head`` multiplies the margins AND the
device-vs-oracle logit error by the same factor (measured: error 0.008 at gain 1, 0.058 at gain 6), so with
random heads ~20 % of the frames of any line stay inside the 2*tolerance band whatever the gain, and no
160-frame line is ever all-safe.  A trained checkpoint has peaked outputs; without one (no network) the heads
are fitted here: the encoder / decoder bodies stay seeded-random ("hardened", kiri_ocr_b200.fixtures), the
oracle computes their features for the fixture lines, and a max-margin linear classifier (squared hinge +
L2, i.e. the smallest weights that reach the margin, so the error amplification is minimal) is solved for
``ctc_head.2`` on chosen frame labels and for ``dec_head`` on the teacher-forced decoder states of the
matching character sequence (with the LM-fusion term of ``lm_head`` as a fixed offset, model.py:480-485).
The result behaves like a trained model on these lines: blank-dominated CTC spikes, a decoder that spells
the same text and stops on EOS.

``make_golden_wide.py`` (build container) runs the fit, checks it with the oracle, runs the UNMODIFIED
reference on the resulting checkpoint and stores weights + reference outputs in ``golden_wide_v1.npz``;
the tests only load that file.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from kiri_ocr_b200 import fixtures as FX

# name -> (state_dict seed, crop rng seed, [(source height, width at H=48)], IMG_W of the reference run per line)
# "wide":   two full-width lines (bucket 640 == the reference's default IMG_W, identical in parity and bucketed mode)
# "wide_b": four lines in the 128 / 256 / 256 / 384 buckets (the reference run with cfg.IMG_W = Wb, core.py:430-431)
WIDE_CASES: Dict[str, dict] = {
    "wide": dict(sd_seed=5, crop_seed=501, lines=[(64, 600), (40, 632)]),
    "wide_b": dict(sd_seed=6, crop_seed=601, lines=[(33, 100), (57, 230), (48, 250), (80, 370)]),
}
MARGIN = 10.0          # fitted top-1 margin (logits / fused log-probs) in fp32


def bucket_of(nw: int) -> int:
    for b in FX.BUCKETS:
        if nw <= b:
            return b
    return FX.BUCKETS[-1]


def wide_crops(name: str) -> Tuple[List[np.ndarray], List[int]]:
    """Seeded crops of a case and the batch width (bucket) of each."""
    case = WIDE_CASES[name]
    rng = np.random.default_rng(case["crop_seed"])
    crops, wbs = [], []
    for h, tw in case["lines"]:
        w = max(1, int(round(tw * h / 48.0)))
        c = FX._draw_line(rng, h, w, inverted=False)
        nw = max(1, int(round(w * (48 / float(h)))))
        crops.append(c)
        wbs.append(bucket_of(nw))
    return crops, wbs


def frame_labels(name: str, line: int, T: int) -> Tuple[np.ndarray, List[int]]:
    """CTC frame labels of a line (runs of 1-3 frames of a character, 1-3 blank frames between) and the
    character sequence (CTC ids, all >= 3: never blank/pad/<unk>)."""
    rng = np.random.default_rng(WIDE_CASES[name]["crop_seed"] * 131 + line)
    lab = np.zeros(T, np.int64)
    seq: List[int] = []
    t = int(rng.integers(1, 3))
    while t < T - 2:
        run = int(rng.integers(1, 4))
        ch = int(rng.integers(3, 204))
        lab[t:t + run] = ch
        seq.append(ch)
        t += run + int(rng.integers(1, 4))
    return lab, seq


def fit_max_margin(feats, labels, n_classes: int, offsets=None, lam: float = 1e-4, steps: int = 6000):
    """Smallest-norm linear head with ``z[y] - z[c] + off[y] - off[c] >= 1`` for every point and class c != y
    (squared hinge + L2 on the weights, bias free), float64 Adam; returns (W [n_classes, D], b [n_classes],
    achieved min margin (>= 1 after the final rescale of the weights), max row norm)."""
    import torch
    F = torch.as_tensor(feats, dtype=torch.float64)
    N, D = F.shape
    y = torch.as_tensor(labels, dtype=torch.long)
    A = torch.cat([F, torch.ones(N, 1, dtype=torch.float64)], 1)
    off = torch.zeros(N, n_classes, dtype=torch.float64) if offsets is None else torch.as_tensor(offsets, dtype=torch.float64)
    W = torch.zeros(D + 1, n_classes, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([W], lr=0.01)
    ar = torch.arange(N)
    for _ in range(steps):
        z = A @ W + off
        viol = (1.0 - (z[ar, y][:, None] - z)).clamp_min(0)
        viol = viol.index_put((ar, y), torch.zeros(N, dtype=torch.float64))
        loss = (viol ** 2).sum(1).mean() + lam * (W[:D] ** 2).sum()
        opt.zero_grad()
        loss.backward()
        opt.step()
    W = W.detach()
    ar_y = (ar, y)

    def min_margin(k):
        z = k * (A @ W) + off
        zy = z[ar_y]
        z[ar_y] = -1e30
        return float((zy - z.max(1).values).min())
    # the hinge target is approached from below: scale the weights (not the fixed offsets) up to reach it exactly
    k_lo, k_hi = 1.0, 1.0
    while min_margin(k_hi) < 1.0 and k_hi < 256:
        k_hi *= 2
    for _ in range(40):
        k = 0.5 * (k_lo + k_hi)
        if min_margin(k) < 1.0:
            k_lo = k
        else:
            k_hi = k
    W = W * k_hi
    return W[:D].T.contiguous(), W[D].contiguous(), min_margin(1.0), float(W[:D].norm(dim=0).max())


def wide_state_dict(gw, name: str):
    """The fixture checkpoint: seeded hardened body + the fitted heads stored in golden_wide_v1.npz."""
    import torch
    from kiri_ocr_b200.config import CFG
    sd = FX.make_state_dict(CFG(), 202, seed=WIDE_CASES[name]["sd_seed"], hardened=True)
    for key, arr in (("ctc_head.2.weight", "ctc_w"), ("ctc_head.2.bias", "ctc_b"), ("dec_head.weight", "dec_w"),
                     ("dec_head.bias", "dec_b")):
        sd[key] = torch.from_numpy(np.ascontiguousarray(gw[f"{name}/{arr}"], dtype=np.float32))
    return sd
