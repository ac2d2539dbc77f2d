"""CPU oracle for the kiri-ocr line-recognition hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm (``/root/reference/kiri_ocr``):
``preprocess.py`` restates Pillow's fixed-point bilinear resample + the crop/invert/pad rules
in numpy integers; ``model.py`` restates the recognizer forward in torch fp32 functional ops
straight from a reference-layout ``state_dict``; ``decode.py`` restates the CTC greedy
collapse and the greedy attention decoder (with a KV cache, which is result-preserving).

Parity pinning: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against *outputs of the reference itself*, generated in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and committed
as ``tests/golden/golden_v1.npz``; ``tests/test_oracle_golden.py`` checks the oracle against
them, and ``tests/test_oracle_vs_pillow.py`` checks the resample against the installed Pillow.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package.  The product path (``kiri_ocr_b200``) never does;
it fails loudly when the CUDA library is missing instead of falling back to this code.
"""
