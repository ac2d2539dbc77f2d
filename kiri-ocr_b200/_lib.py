"""ctypes binding of libkiri_b200.so (the C ABI declared in include/kiri_b200.h).

There is NO fallback: if the shared library is missing or the device is not a B200-class GPU,
importing callers get a RuntimeError that says how to build it.  PyTorch only provides device
memory (``tensor.data_ptr()``) and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# KIRI_B200_LIB: load another build of the same ABI (the checked build of the soak test: `make -C csrc checked`)
LIB_PATH = os.environ.get("KIRI_B200_LIB") or os.path.join(_PKG_DIR, "libkiri_b200.so")
CHECKED_LIB_PATH = os.path.join(_PKG_DIR, "libkiri_b200_checked.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
KIRI_MAX_LAYERS = 8

DTYPE_F32, DTYPE_BF16 = 0, 1
EPI_BIAS_BF16, EPI_BIAS_SILU_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_BIAS_F32, EPI_BIAS_RESID_LN, EPI_CTC_STATS = range(7)

vp, fp, ip = C.c_void_p, C.c_void_p, C.c_void_p      # all device pointers travel as integers


class KiriCropDesc(C.Structure):
    _fields_ = [("src_offset", C.c_int64), ("out_offset", C.c_int64), ("pitch", C.c_int32), ("w", C.c_int32),
                ("h", C.c_int32), ("nw", C.c_int32), ("Wb", C.c_int32), ("strip_w", C.c_int32), ("flags", C.c_int32),
                ("reserved", C.c_int32)]


CROP_NO_INVERT = 1


class KiriDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "img_h", "enc_dim", "enc_layers", "enc_heads", "enc_ff", "dec_dim", "dec_layers", "dec_heads",
        "dec_ff", "ctc_classes", "dec_vocab", "max_pos", "max_t", "has_dec_pos")]


class KiriEncLayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("wqkv", "bqkv", "wo", "bo", "w1", "b1", "w2", "b2",
                                  "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


class KiriDecLayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("wqkv", "bqkv", "wo", "bo", "wcq", "bcq", "wco", "bco", "w1", "b1", "w2", "b2",
                                  "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b")]


class KiriWeights(C.Structure):
    _fields_ = ([(n, vp) for n in ("conv1_w_host", "conv1_b_host", "conv2_w", "conv2_b", "conv3_w", "conv3_b",
                                   "conv4_w", "conv4_b", "pos_table", "enc_ln_in_g", "enc_ln_in_b")]
                + [("enc", KiriEncLayerWeights * KIRI_MAX_LAYERS)]
                + [(n, vp) for n in ("enc_ln_g", "enc_ln_b", "ctc_ln_g", "ctc_ln_b", "ctc_w", "ctc_b",
                                     "crosskv_w", "crosskv_b", "dec_emb", "dec_pe")]
                + [("dec", KiriDecLayerWeights * KIRI_MAX_LAYERS)]
                + [(n, vp) for n in ("dec_ln_g", "dec_ln_b", "heads_w", "heads_b")])


class KiriGroup(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("n_lines", C.c_int32), ("Wb", C.c_int32)]


class KiriDecodeParams(C.Structure):
    _fields_ = [("lm_alpha", C.c_float), ("eos_bias", C.c_float), ("eos_boost", C.c_float),
                ("eos_bias_until_len", C.c_int32), ("rep_last", C.c_float), ("rep_bigram", C.c_float),
                ("rep_trigram", C.c_float), ("unk_penalty", C.c_float), ("unk_id", C.c_int32),
                ("len_ratio", C.c_double), ("len_pad", C.c_int32), ("mem_ratio", C.c_double),
                ("max_dec_len", C.c_int32), ("select_raw", C.c_int32)]


_SIGS = {
    "kiri_last_error": (C.c_char_p, []),
    "kiri_version": (C.c_int, []),
    "kiri_abi_sizes": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "kiri_device_ok": (C.c_int, []),
    "kiri_profile_begin": (C.c_int, []),
    "kiri_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]),
    "kiri_preprocess_smem_bytes": (C.c_int, [C.c_int] * 6),
    "kiri_preprocess_pack": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "kiri_bgr_to_gray": (C.c_int, [vp, C.c_longlong, vp, vp]),
    "kiri_conv1": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kiri_conv1_multi": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, vp, vp,
                                   C.c_int, vp]),
    "kiri_conv1_tc_multi": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, vp, vp,
                                      C.c_int, vp]),
    "kiri_pool_pos_ln_multi": (C.c_int, [C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, vp, C.c_int, C.c_int,
                                         vp, vp, vp, vp, vp, vp, vp]),
    "kiri_conv3x3_bf16": (C.c_int, [vp, vp, vp] + [C.c_int] * 7 + [vp, C.c_int, vp]),
    "kiri_gemm_bf16": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    "kiri_gemm_ref": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kiri_pool_pos_ln": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "kiri_layernorm": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
    "kiri_encoder_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kiri_pack_records": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "kiri_encoder_block": (C.c_int, [vp] * 13 + [C.c_int, C.c_int, vp]),
    "kiri_encoder_block_soak": (C.c_int, [vp] * 13 + [C.c_int, C.c_int, C.c_int, vp]),
    "kiri_encoder_attention_multi": (C.c_int, [vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, vp, vp]),
    "kiri_ctc_greedy": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    "kiri_ctc_greedy_multi": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    "kiri_create": (C.c_int, [C.POINTER(KiriDims), C.POINTER(KiriWeights), C.POINTER(vp)]),
    "kiri_destroy": (None, [vp]),
    "kiri_encode_workspace_bytes": (C.c_size_t, [vp, C.c_int, C.c_int, C.c_int]),
    "kiri_encode": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]),
    "kiri_encode_multi_workspace_bytes": (C.c_size_t, [vp, C.POINTER(KiriGroup), C.c_int, C.c_int]),
    "kiri_encode_multi": (C.c_int, [vp, C.POINTER(KiriGroup), C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp, vp]),
    "kiri_ctc_collapse_multi": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]),
    "kiri_decode_workspace_bytes": (C.c_size_t, [vp, C.c_int, C.c_int, C.c_int]),
    "kiri_decode_multi_workspace_bytes": (C.c_size_t, [vp, C.c_int, C.c_longlong, C.c_int]),
    "kiri_decode_greedy_multi": (C.c_int, [vp, vp, C.c_longlong, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(KiriDecodeParams),
                                           vp, C.c_size_t, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp, C.c_int, vp]),
    "kiri_decode_beam_workspace_bytes": (C.c_size_t, [vp, C.c_int, C.c_longlong, C.c_int, C.c_int]),
    "kiri_decode_beam_multi": (C.c_int, [vp, vp, C.c_longlong, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.POINTER(KiriDecodeParams), vp, C.c_size_t, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "kiri_ctc_align_score": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int,
                                       C.c_int, vp, vp]),
    "kiri_decode_greedy": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(KiriDecodeParams), vp,
                                     C.c_size_t, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), C.c_int, vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib: Optional[C.CDLL] = None


class KiriError(RuntimeError):
    pass


def build_library(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4), "all", "checked"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise KiriError("building libkiri_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (no compute is issued).  Raises KiriError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KiriError(
            f"{LIB_PATH} is missing: the B200 CUDA library is not built and there is no CPU fallback. "
            f"Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C {CSRC_DIR}`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_int * 6)()
    lib.kiri_abi_sizes(sizes, 6)
    mine = [C.sizeof(t) for t in (KiriCropDesc, KiriDims, KiriWeights, KiriGroup, KiriDecodeParams, KiriEncLayerWeights)]
    if list(sizes) != mine:
        raise KiriError(f"{LIB_PATH} was built against another include/kiri_b200.h (struct sizes {list(sizes)} != {mine}): "
                        f"rebuild it with `make -C {CSRC_DIR}`")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().kiri_last_error().decode("utf-8", "replace")
        raise KiriError(f"{what or 'libkiri_b200'} failed ({rc}): {msg}")


def require_device() -> None:
    """Fail loudly unless a CUDA device of compute capability 10.x is current."""
    import torch
    if not torch.cuda.is_available():
        raise KiriError("kiri_ocr_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU path")
    if not load().kiri_device_ok():
        raise KiriError("kiri_ocr_b200 kernels are built for sm_100a only; the current device is not compute capability 10.x")


PROFILE_STAGES = ("conv1", "conv2", "conv3", "conv4", "pool_ln", "qkv", "attention", "out_proj", "ff1", "ff2",
                  "ln_final", "ctc_head", "dec_crosskv", "dec_step", "preprocess", "ctc_greedy")


def ptr(t) -> int:
    """Device (or host) address of a tensor / None -> 0."""
    return 0 if t is None else t.data_ptr()


# The engine pins the stream handle while it runs a batch on its own stream: torch.cuda.current_stream() costs ~4 us and a
# 256-line batch asks for it in front of every one of its ~15 library calls.
_stream_override: Optional[int] = None


def stream_ptr() -> int:
    if _stream_override is not None:
        return _stream_override
    import torch
    return torch.cuda.current_stream().cuda_stream


class pinned_stream:
    """``with pinned_stream(stream.cuda_stream):`` - stream_ptr() returns that handle inside (the caller guarantees that the
    stream IS the current one for the duration)."""

    def __init__(self, handle: int):
        self.handle, self.prev = handle, None

    def __enter__(self):
        global _stream_override
        self.prev, _stream_override = _stream_override, self.handle
        return self

    def __exit__(self, *exc):
        global _stream_override
        _stream_override = self.prev
        return False
