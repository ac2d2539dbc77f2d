"""Stated bf16 tolerances of the device path against the fp32 oracle (north_star: "within a stated bf16 tolerance").

The device computes with bf16 operands, fp32 accumulation and an fp32 residual stream.  Its error at the encoder
output is bounded in absolute terms (LayerNorm-ed activations are O(1)); behind an output head it is LINEAR in that
head's weights, so the logit / log-prob tolerances are stated per unit of the head's largest row norm.  Measured on
the B200 (profiles/r02_parity_report.json): mem max-abs 0.012-0.014, rel-L2 0.0027; CTC logits 0.0168 x row norm
(0.058 at the "hard" fixture's row norm 3.5, 0.008 at default init's 0.58); decoder step log-probs 0.009 x
(|dec_head| + alpha |lm_head|) (0.042 on "hard").  The tolerances are 2x the measured errors:

    MEM_ATOL 0.03, MEM_RTOL 0.006; logit_tol = 0.033 |W_ctc|  (= 0.12 on "hard"); dec_tol = 0.02 (|W_dec| + alpha |W_lm|)
    (= 0.10 on "hard").

A frame / step is MARGIN-SAFE when the oracle's top-1 margin exceeds 2 x tolerance; ids must be bit-exact there.
"""
import torch

MEM_ATOL, MEM_RTOL = 0.03, 0.006
LOGIT_RTOL_PER_NORM = 0.033
DEC_RTOL_PER_NORM = 0.02


def _rownorm(w: torch.Tensor) -> float:
    return float(w.float().norm(dim=1).max())


def logit_tol(sd) -> float:
    return LOGIT_RTOL_PER_NORM * _rownorm(sd["ctc_head.2.weight"])


def dec_tol(sd, lm_alpha: float = 0.35) -> float:
    n = _rownorm(sd["dec_head.weight"])
    if "lm_head.weight" in sd:
        n += lm_alpha * _rownorm(sd["lm_head.weight"])
    return DEC_RTOL_PER_NORM * n
