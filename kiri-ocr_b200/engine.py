"""BatchedRecognizer — host orchestration of the B200 line-recognition path.

Takes line crops (or pages + boxes), runs preprocess -> stem -> encoder -> CTC head -> CTC greedy
("ctc") or the greedy attention decoder ("decoder") on the device through ``libkiri_b200.so``, and
returns the reference's ``(text, confidence)`` per line.  It replaces the per-line loop body of
``OCR.process_document`` (kiri_ocr/core.py:770-776: ``_preprocess_region`` + ``recognize_region``).

Width modes:
  * ``"parity"``   — every line is padded to ``cfg.IMG_W`` (640) exactly like the reference;
  * ``"bucketed"`` — lines are grouped by the smallest bucket in {128,...,640} that holds their
    resized width; a group of width Wb equals the reference run with ``cfg.IMG_W = Wb``;
  * ``"masked"``   — bucketed + per-line key mask in self-attention (no reference equivalent).
PyTorch is used for device memory, pinned staging and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
import os as _os
import time as _time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .config import CFG, CharTokenizer
from .weights import PackedWeights

BUCKETS = (128, 256, 384, 512, 640)
DESC_DTYPE = np.dtype([("src_offset", "<i8"), ("out_offset", "<i8"), ("pitch", "<i4"), ("w", "<i4"), ("h", "<i4"),
                       ("nw", "<i4"), ("Wb", "<i4"), ("strip_w", "<i4"), ("flags", "<i4"), ("reserved", "<i4")])
assert DESC_DTYPE.itemsize == C.sizeof(_lib.KiriCropDesc)
PRE_SMEM_CAP = 56 * 1024           # four preprocessing CTAs per SM
PRE_STRIP = 128                    # output columns per preprocessing CTA


@dataclass(slots=True)
class LineResult:
    text: str
    confidence: float
    ctc_confidence: float
    ids: np.ndarray                 # collapsed CTC ids ("ctc") or decoder ids ("decoder")
    step_logp: Optional[np.ndarray] = None
    step_prob: Optional[np.ndarray] = None
    frame_ids: Optional[np.ndarray] = None
    frame_prob: Optional[np.ndarray] = None
    len_est: int = 0                # CTC length estimate that bounds the decoder (model.py:416-425)


class _Raw:
    """A raw (device-visible) address standing in for a tensor in the low-level calls."""
    __slots__ = ("p",)

    def __init__(self, p: int):
        self.p = p

    def data_ptr(self) -> int:
        return self.p


class LineError:
    """A region that could not be recognised (malformed box, per-line failure): the document loops drop it or
    report it as an ``error`` result / chunk like the reference's per-region ``try/except``
    (kiri_ocr/core.py:771-791, 873-885, 1011-1026)."""
    __slots__ = ("message",)

    def __init__(self, message: str):
        self.message = message

    def __repr__(self):
        return f"LineError({self.message!r})"


def target_widths(w: np.ndarray, h: np.ndarray, img_h: int) -> np.ndarray:
    """``max(1, int(round(iw * (img_h / float(ih)))))`` (model.py:321-322), vectorised; numpy's
    rint is round-half-to-even like Python's round."""
    scale = float(img_h) / h.astype(np.float64)
    return np.maximum(1, np.rint(w.astype(np.float64) * scale)).astype(np.int64)


def _pre_smem_parts(w, h, nw, img_h, Wb, strip):
    """numpy mirror of kiri_preprocess_smem_bytes (csrc/preprocess.cu), split into the part that does not depend on the
    number of staged source rows and the bytes per staged row: smem(rows) = fixed + per_row * rows."""
    wout = np.minimum(nw, Wb)
    ws = np.minimum(strip, wout)
    hs = w / nw
    vs = h / img_h
    hs1 = np.maximum(hs, 1.0)
    ksh = np.where(nw != w, np.ceil(hs1).astype(np.int64) * 2 + 1, 1)
    ksv = np.where(h != img_h, np.ceil(np.maximum(vs, 1.0)).astype(np.int64) * 2 + 1, 1)
    fixed = (((img_h * 4) * ksv + 15) & ~15) + ((img_h * 4 + 15) & ~15) + ((ws * ksh * 4 + 15) & ~15) + ((ws * 4 + 15) & ~15) \
        + ((h * ((ws + 3) & ~3) + 15) & ~15)
    span = np.ceil(hs1 * ws).astype(np.int64) + 2 * ksh + 8
    per_row = ((span + 4 + 15) & ~15) + 16
    return fixed, per_row


def _pre_smem(w, h, nw, img_h, Wb, strip, rows=8):
    """smem bytes of kiri_preprocess_pack for ``rows`` staged source rows (8 = the minimum the C helper reports)."""
    fixed, per_row = _pre_smem_parts(w, h, nw, img_h, Wb, strip)
    return fixed + per_row * rows


def plan_groups(entries: np.ndarray, cfg: CFG, width_mode: str = "parity"
                ) -> Dict[int, Tuple[np.ndarray, np.ndarray, int]]:
    """Group lines by batch width; returns {Wb: (line indices, descriptors, smem bytes, max strips)}."""
    img_h = cfg.IMG_H
    w, h = entries[:, 2], entries[:, 3]
    nw = target_widths(w, h, img_h)
    if width_mode == "parity":
        wb = np.full(len(entries), cfg.IMG_W, np.int64)
    else:
        bk = np.array(tuple(b for b in BUCKETS if b <= cfg.IMG_W) or (cfg.IMG_W,), np.int64)
        wb = bk[np.minimum(np.searchsorted(bk, np.minimum(nw, bk[-1])), len(bk) - 1)]
    # everything per line is computed once for the whole batch; the per-group work is slicing only
    wout = np.minimum(nw, wb)
    strip = np.minimum(wout, PRE_STRIP)
    fixed, per_row = _pre_smem_parts(w, h, nw, img_h, wb, strip)
    need = fixed + per_row * 8
    for _ in range(6):                                       # halve strips until they fit
        big = need > PRE_SMEM_CAP
        if not big.any():
            break
        strip = np.where(big & (strip > 32), np.maximum(32, (strip // 2 + 31) // 32 * 32), strip)
        fixed, per_row = _pre_smem_parts(w, h, nw, img_h, wb, strip)
        need = fixed + per_row * 8
    # more shared memory than the minimum lets a CTA stage every source row of its strip at once
    need = np.maximum(need, np.minimum(fixed + per_row * h, PRE_SMEM_CAP))
    nstr = (wout + strip - 1) // strip
    d_all = np.zeros(len(entries), DESC_DTYPE)
    d_all["src_offset"], d_all["pitch"], d_all["w"], d_all["h"] = entries[:, 0], entries[:, 1], w, h
    d_all["nw"], d_all["strip_w"], d_all["Wb"] = nw, strip, wb
    if entries.shape[1] > 4:                                 # optional 5th column: KiriCropDesc.flags (CROP_NO_INVERT)
        d_all["flags"] = entries[:, 4]
    order = np.argsort(wb, kind="stable")
    wbs = wb[order]
    cuts = np.nonzero(np.diff(wbs))[0] + 1
    groups = {}
    for lo, hi in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [len(order)]])):
        idx = order[lo:hi]
        d = d_all[idx]
        d["out_offset"] = np.arange(hi - lo, dtype=np.int64) * (img_h * int(wbs[lo]))   # inside the group's planes
        groups[int(wbs[lo])] = (idx, d, int(need[idx].max()), int(nstr[idx].max()))
    return groups


def decode_slot_table(B: int, n0: Optional[int] = None, lines_per_cluster: int = 16, resident_clusters: int = 15) -> np.ndarray:
    """Decode slot -> rank of the line in decreasing ``len_est`` order, -1 = empty slot (int64 [16 * clusters]).

    Sixteen consecutive slots are one thread-block cluster of the persistent decode kernel.  Cluster k holds
    ``min(16, n0 + k)`` lines: a decode step costs a cluster a fixed part plus a part per live line
    (profiles/r02_dec_fused_phase_cycles.txt), the decode ends when the cluster with the longest lines does, and
    ``resident_clusters`` clusters are co-resident on a B200 - so the clusters that hold the longest lines are kept small:
    the smallest n0 >= 4 that still needs at most ``resident_clusters`` clusters, and n0 = 9 (the measured optimum at 256
    lines, tools/runs/r2_run28.sh) when no n0 does.  The table depends on B only: building it never waits for the device."""
    def caps(first):
        out, left, k = [], B, 0
        while left > 0:
            c = min(lines_per_cluster, first + k, left)
            out.append(c)
            left -= c
            k += 1
        return out
    if n0 is not None:
        cs = caps(max(1, int(n0)))
    else:
        cs = next((c for c in (caps(f) for f in range(4, lines_per_cluster + 1)) if len(c) <= resident_clusters), None) or caps(9)
    t = np.full(lines_per_cluster * len(cs), -1, np.int64)
    r = 0
    for k, c in enumerate(cs):
        t[lines_per_cluster * k:lines_per_cluster * k + c] = np.arange(r, r + c)
        r += c
    return t


class BatchedRecognizer:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: CFG, tokenizer: CharTokenizer,
                 device: str = "cuda", width_mode: str = "parity", stem_chunk: int = 64):
        _lib.require_device()
        self.lib = _lib.load()
        self.cfg, self.tok = cfg, tokenizer
        self.device = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        if self.device.index != torch.cuda.current_device():
            # the C ABI works on the CURRENT device (handle creation, cudaMalloc of the packed decoder weights,
            # every launch); silently mixing devices would read another GPU's pointers
            raise _lib.KiriError(f"BatchedRecognizer(device={device!r}): make it the current device first "
                                 f"(torch.cuda.set_device({self.device.index})); the current device is "
                                 f"cuda:{torch.cuda.current_device()}")
        if width_mode not in ("parity", "bucketed", "masked"):
            raise ValueError(f"unknown width_mode {width_mode!r}")
        self.width_mode = width_mode
        self.stem_chunk = stem_chunk
        self.pw = PackedWeights(state_dict, cfg, tokenizer.vocab_size, self.device)
        h = C.c_void_p()
        _lib.check(self.lib.kiri_create(C.byref(self.pw.dims), C.byref(self.pw.struct), C.byref(h)), "kiri_create")
        self.handle = h
        # the engine runs on its own (capturable) stream; public calls fence it against the caller's stream
        self.stream = torch.cuda.Stream(device=self.device)
        self._copy_stream = torch.cuda.Stream(device=self.device)     # uploads overlap the previous batch
        self._slot = 0                                                # ping-pong staging slot of submit()
        self._src_free = [None, None]
        self._h2d_done = [None, None]                                 # H2D of the slot's pinned source buffer
        self._ws: Optional[torch.Tensor] = None
        self._dws: Optional[torch.Tensor] = None
        self._build_tables()
        self.launches = 0           # kernels launched by this engine (for bench's gpu_launches)
        # per encoder layer: QKV GEMM, attention, fused tail (csrc/encoder_block.cu) — or QKV, attention and the
        # three GEMMs the tail replaces when it is switched off / not applicable
        fused = _os.environ.get("KIRI_NO_FUSED_BLOCK") is None and cfg.ENC_FF % 256 == 0 and cfg.ENC_FF <= 1024
        self._layer_launches = 3 if fused else 5

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.kiri_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ host-side planning
    def buckets(self) -> Tuple[int, ...]:
        return tuple(b for b in BUCKETS if b <= self.cfg.IMG_W) or (self.cfg.IMG_W,)

    @staticmethod
    def pack_crops(crops: Sequence[np.ndarray]) -> Tuple[torch.Tensor, np.ndarray]:
        """Concatenate uint8 crops into one pinned buffer; returns (buffer, entries[n,4] =
        offset, pitch, w, h)."""
        sizes = np.array([c.size for c in crops], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if len(crops) else np.zeros(0, np.int64)
        buf = torch.empty(int(sizes.sum()) + 16, dtype=torch.uint8).pin_memory()
        nb = buf.numpy()
        ent = np.zeros((len(crops), 4), np.int64)
        for i, c in enumerate(crops):
            if c.dtype != np.uint8 or c.ndim != 2:
                raise ValueError("crops must be 2-D uint8 arrays")
            nb[offs[i]:offs[i] + c.size] = np.ascontiguousarray(c).reshape(-1)
            ent[i] = (offs[i], c.shape[1], c.shape[1], c.shape[0])
        return buf, ent

    @staticmethod
    def boxes_to_entries(page_shape: Tuple[int, int], boxes: Sequence[Sequence[int]], page_offset: int = 0,
                         extra_padding: int = 5) -> Tuple[np.ndarray, np.ndarray]:
        """core.py:506-517 for every box: clamp-pad by 5 px; returns (entries, valid mask)."""
        H, W = page_shape
        b = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
        x1 = np.maximum(0, b[:, 0] - extra_padding)
        y1 = np.maximum(0, b[:, 1] - extra_padding)
        x2 = np.minimum(W, b[:, 0] + b[:, 2] + extra_padding)
        y2 = np.minimum(H, b[:, 1] + b[:, 3] + extra_padding)
        valid = (x2 > x1) & (y2 > y1)
        ent = np.stack([page_offset + y1 * W + x1, np.full_like(x1, W), x2 - x1, y2 - y1], axis=1)
        return ent, valid

    def plan(self, entries: np.ndarray) -> Dict[int, Tuple[np.ndarray, np.ndarray, int, int]]:
        return plan_groups(entries, self.cfg, self.width_mode)

    # ------------------------------------------------------------------ device stages
    def _workspace(self, nbytes: int, which: str = "_ws") -> torch.Tensor:
        cur = getattr(self, which)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            setattr(self, which, cur)
        return cur

    def _pinned(self, name: str, n: int, dtype) -> torch.Tensor:
        """Persistent pinned host staging buffer of at least n elements."""
        cur = getattr(self, name, None)
        if cur is None or cur.numel() < n:
            cur = torch.empty(max(n, 1024), dtype=dtype).pin_memory()
            setattr(self, name, cur)
        return cur

    def _device(self, name: str, n: int, dtype) -> torch.Tensor:
        cur = getattr(self, name, None)
        if cur is None or cur.numel() < n:
            cur = torch.empty(max(n, 1024), dtype=dtype, device=self.device)
            setattr(self, name, cur)
        return cur

    def _build_tables(self):
        """id -> text lookup lists (the tokenizer's tables padded to the device's class counts, so ids of the zero
        padded head columns map to '')."""
        tok = self.tok
        self._ctc_table = list(tok.ctc_text) + [""] * (self.pw.Cp + 1 - len(tok.ctc_text))
        self._dec_table = list(tok.dec_text) + [""] * (self.pw.Vp + 1 - len(tok.dec_text))

    def preprocess(self, src_dev: torch.Tensor, descs: np.ndarray, Wb: int, smem: int, n_strips: int,
                   want_norm: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        n = len(descs)
        dd = torch.from_numpy(descs.view(np.uint8).reshape(-1).copy()).to(self.device)
        planes = torch.empty((n, self.cfg.IMG_H, Wb), dtype=torch.uint8, device=self.device)
        norm = torch.empty((n, self.cfg.IMG_H, Wb), dtype=torch.bfloat16, device=self.device) if want_norm else None
        sums = torch.empty(2 * n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.kiri_preprocess_pack(src_dev.data_ptr(), dd.data_ptr(), n, self.cfg.IMG_H, smem, n_strips,
                                                 planes.data_ptr(), _lib.ptr(norm), sums.data_ptr(), _lib.stream_ptr()),
                   "kiri_preprocess_pack")
        self.launches += 2
        return planes, norm

    def encode(self, planes: torch.Tensor, want_mem_f32: bool = False, want_tokens: bool = False,
               kv_len: Optional[torch.Tensor] = None, want_logits: bool = True):
        """[B, IMG_H, Wb] uint8 -> dict(mem_bf16, logits [B,T,Cp], mem_f32?, tokens?)."""
        B, H, Wb = planes.shape
        T, D = Wb // 4, self.cfg.ENC_DIM
        need = self.lib.kiri_encode_workspace_bytes(self.handle, B, Wb, self.stem_chunk)
        ws = self._workspace(need)
        out = {"mem_bf16": torch.empty((B * T, D), dtype=torch.bfloat16, device=self.device)}
        if want_logits:
            out["logits"] = torch.empty((B, T, self.pw.Cp), dtype=torch.float32, device=self.device)
        if want_mem_f32:
            out["mem_f32"] = torch.empty((B, T, D), dtype=torch.float32, device=self.device)
        if want_tokens:
            out["tokens"] = torch.empty((B, T, D), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.kiri_encode(self.handle, planes.data_ptr(), B, Wb, self.stem_chunk, ws.data_ptr(), need,
                                        _lib.ptr(out.get("mem_f32")), out["mem_bf16"].data_ptr(),
                                        _lib.ptr(out.get("logits")), _lib.ptr(out.get("tokens")), _lib.ptr(kv_len),
                                        _lib.stream_ptr()), "kiri_encode")
        self.launches += self._stem_launches([(B, Wb)]) + 1 + self._layer_launches * self.pw.enc_layers + 1 + (1 if want_logits else 0)
        return out

    def _stem_sub_batch(self, B: int, Wb: int) -> int:
        """Mirror of stem_sub_batch() in csrc/api.cu (only used to count launches)."""
        sc = self.stem_chunk
        if sc <= 0:
            return B
        cap = max(sc, sc * 640 // Wb)
        if cap >= B:
            return B
        n = -(-B // cap)
        return -(-B // n)

    @staticmethod
    def _conv_nseg(OH: int, OW: int) -> int:
        """Tile form of a conv problem (mirror of prep_problem in csrc/gemm_tc.cu): 4 = four 32-pixel segments."""
        best = min(((OW + S - 1) // S * S) * ((OH + R - 1) // R * R) / (OW * OH) for R, S in ((1, 128), (2, 64), (4, 32)))
        return 4 if best > 1.25 else 1

    def _stem_launches(self, groups) -> int:
        """Kernel launches of the stem for width groups [(lines, Wb)]: per round one conv1 and one launch per conv
        layer and tile form (csrc/api.cu; only used for the bench's launch count)."""
        H = self.cfg.IMG_H
        subs = [self._stem_sub_batch(B, Wb) for B, Wb in groups]
        rounds = max(-(-B // sc) for (B, _), sc in zip(groups, subs))
        n = 0
        for r in range(rounds):
            live = [Wb for (B, Wb), sc in zip(groups, subs) if r * sc < B]
            n += 1                                            # conv1 of all groups
            for OH, div in ((H // 2, 2), (H // 4, 4), (H // 8, 4)):
                n += len({self._conv_nseg(OH, Wb // div) for Wb in live})
        return n

    def encode_multi(self, planes_list: Sequence[torch.Tensor], want_mem_f32: bool = False, want_tokens: bool = False,
                     kv_len: Optional[torch.Tensor] = None, want_logits: bool = True, slot: Optional[str] = None,
                     stats: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, want_mem: bool = True):
        """Several width groups ([B_g, IMG_H, Wb_g] uint8 each) in one call: per-group stems, ONE pass
        of the encoder / CTC head over the concatenated token stream.  Returns the outputs token-major
        ([M, ...], M = sum B_g * Wb_g / 4) plus ``rows`` = [(row0, B_g, T_g)] per group.  ``stats`` = (frame_ids int32
        [M], frame_prob fp32 [M]): the CTC head's epilogue takes the per-token arg-max and its probability; with
        ``want_logits=False`` the logits then never reach HBM; ``want_mem=False`` (method "ctc": nothing reads the
        encoder memory) skips the bf16 ``mem`` store of the final LayerNorm as well."""
        D = self.cfg.ENC_DIM
        n = len(planes_list)
        garr = (_lib.KiriGroup * n)()
        rows, M = [], 0
        for i, pl in enumerate(planes_list):
            B, H, Wb = pl.shape
            garr[i].planes, garr[i].n_lines, garr[i].Wb = pl.data_ptr(), B, Wb
            rows.append((M, B, Wb // 4))
            M += B * (Wb // 4)
        need = self.lib.kiri_encode_multi_workspace_bytes(self.handle, garr, n, self.stem_chunk)
        ws = self._workspace(need)
        if slot is None:
            out = {"rows": rows}
            if want_mem:
                out["mem_bf16"] = torch.empty((M, D), dtype=torch.bfloat16, device=self.device)
            if want_logits:
                out["logits"] = torch.empty((M, self.pw.Cp), dtype=torch.float32, device=self.device)
        else:
            # submit(): outputs live in the ticket's ping-pong slot (no allocator / tensor-map-cache churn per call)
            out = {"rows": rows}
            if want_mem:
                out["mem_bf16"] = self._device("_mem" + slot, M * D, torch.bfloat16)[:M * D].view(M, D)
            if want_logits:
                out["logits"] = self._device("_logits" + slot, M * self.pw.Cp, torch.float32)[:M * self.pw.Cp].view(M, self.pw.Cp)
        if want_mem_f32:
            out["mem_f32"] = torch.empty((M, D), dtype=torch.float32, device=self.device)
        if want_tokens:
            out["tokens"] = torch.empty((M, D), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.kiri_encode_multi(self.handle, garr, n, self.stem_chunk, ws.data_ptr(), need,
                                              _lib.ptr(out.get("mem_f32")), _lib.ptr(out.get("mem_bf16")),
                                              _lib.ptr(out.get("logits")), _lib.ptr(out.get("tokens")), _lib.ptr(kv_len),
                                              _lib.ptr(stats[0]) if stats else 0, _lib.ptr(stats[1]) if stats else 0,
                                              _lib.stream_ptr()), "kiri_encode_multi")
        self.launches += self._stem_launches([(B, 4 * T) for _, B, T in rows]) + 1             # stem + pool
        self.launches += self._layer_launches * self.pw.enc_layers + 1 + (1 if (want_logits or stats) else 0)
        return out

    def _slot_table(self, B: int) -> torch.Tensor:
        """Decode slot -> rank of the line in decreasing len_est order, -1 = empty slot (device int64, cached per B);
        see ``decode_slot_table``."""
        if not hasattr(self, "_slot_tables"):
            self._slot_tables = {}
        tab = self._slot_tables.get(B)
        if tab is None:
            fixed = _os.environ.get("KIRI_DEC_SLOTS_N0")
            tab = torch.from_numpy(decode_slot_table(B, int(fixed) if fixed else None)).to(self.device)
            self._slot_tables[B] = tab
        return tab

    def decode_greedy_multi(self, mem_bf16: torch.Tensor, mem_row0: torch.Tensor, mem_len: torch.Tensor,
                            len_est: torch.Tensor, Lmax: int, max_T: int, select_raw: bool = False,
                            forced: Optional[torch.Tensor] = None, want_steps: bool = False, out=None,
                            progress_ptr: int = 0, publish: bool = False):
        """One persistent decode over every line of every group (concatenated token stream).  ``out`` tensors (or
        raw addresses, for mapped pinned host memory) receive the results; ``progress_ptr`` / ``publish``: see
        kiri_decode_greedy_multi (live streaming)."""
        p = self.decode_params(select_raw)
        B, M = int(len_est.numel()), int(mem_bf16.shape[0])
        # longest lines first: sixteen consecutive slots share a cluster, and the clusters that hold the longest lines
        # get fewer of them (see kiri_b200.h); the slot table depends on B only, so nothing here waits for the device
        table = self._slot_table(B)
        n_slots = int(table.numel())
        need = self.lib.kiri_decode_multi_workspace_bytes(self.handle, n_slots, M, Lmax)
        ws = self._workspace(need, "_dws")
        order = torch.argsort(len_est, descending=True, stable=True)
        perm = torch.where(table >= 0, order[table.clamp(min=0)], table).to(torch.int32)
        if out is not None:
            ids, n_out, sum_lp, slp, spr = out
        else:
            ids = torch.zeros((B, Lmax), dtype=torch.int32, device=self.device)
            n_out = torch.zeros(B, dtype=torch.int32, device=self.device)
            sum_lp = torch.zeros(B, dtype=torch.float32, device=self.device)
            slp = torch.zeros((B, Lmax), dtype=torch.float32, device=self.device) if want_steps else None
            spr = torch.zeros((B, Lmax), dtype=torch.float32, device=self.device) if want_steps else None
        _lib.check(self.lib.kiri_decode_greedy_multi(self.handle, mem_bf16.data_ptr(), M, mem_row0.data_ptr(),
                                                     mem_len.data_ptr(), max_T, len_est.data_ptr(), perm.data_ptr(), n_slots, B,
                                                     Lmax, C.byref(p),
                                                     ws.data_ptr(), need, ids.data_ptr(), n_out.data_ptr(),
                                                     sum_lp.data_ptr(), _lib.ptr(slp), _lib.ptr(spr), _lib.ptr(forced),
                                                     None, progress_ptr, int(publish), _lib.stream_ptr()),
                   "kiri_decode_greedy_multi")
        self.launches += 2          # cross-K/V GEMM + the persistent decode kernel
        return ids, n_out, sum_lp, slp, spr

    def ctc_greedy(self, logits: torch.Tensor, want_frames: bool = False):
        B, T, Cp = logits.shape
        ids = torch.empty((B, T), dtype=torch.int32, device=self.device)
        n_ids = torch.empty(B, dtype=torch.int32, device=self.device)
        conf = torch.empty(B, dtype=torch.float32, device=self.device)
        fids = torch.empty((B, T), dtype=torch.int32, device=self.device) if want_frames else None
        fprob = torch.empty((B, T), dtype=torch.float32, device=self.device) if want_frames else None
        _lib.check(self.lib.kiri_ctc_greedy(logits.data_ptr(), _lib.DTYPE_F32, B, T, self.pw.C, Cp, ids.data_ptr(),
                                            n_ids.data_ptr(), conf.data_ptr(), _lib.ptr(fids), _lib.ptr(fprob),
                                            _lib.stream_ptr()), "kiri_ctc_greedy")
        self.launches += 1
        return ids, n_ids, conf, fids, fprob

    def decode_params(self, select_raw: bool = False) -> "_lib.KiriDecodeParams":
        cfg = self.cfg
        p = _lib.KiriDecodeParams()
        fuse = cfg.USE_LM and cfg.USE_LM_FUSION_EVAL and self.pw.has_lm
        p.lm_alpha = cfg.LM_FUSION_ALPHA if fuse else 0.0
        p.eos_bias, p.eos_boost, p.eos_bias_until_len = cfg.EOS_LOGP_BIAS, cfg.EOS_LOGP_BOOST, cfg.EOS_BIAS_UNTIL_LEN
        p.rep_last, p.rep_bigram, p.rep_trigram = cfg.REPEAT_LAST_PENALTY, cfg.REPEAT_BIGRAM_PENALTY, cfg.REPEAT_TRIGRAM_PENALTY
        p.unk_penalty, p.unk_id = cfg.UNK_LOGP_PENALTY, self.tok.unk_id + self.tok.dec_offset
        p.len_ratio, p.len_pad, p.mem_ratio, p.max_dec_len = cfg.DEC_MAX_LEN_RATIO, cfg.DEC_MAX_LEN_PAD, cfg.MEM_MAX_LEN_RATIO, cfg.MAX_DEC_LEN
        p.select_raw = int(select_raw)
        return p

    def max_steps_bound(self, len_est_max: int, T: int) -> int:
        cfg = self.cfg
        a = min(cfg.MAX_DEC_LEN, int(len_est_max * cfg.DEC_MAX_LEN_RATIO) + cfg.DEC_MAX_LEN_PAD)
        b = min(cfg.MAX_DEC_LEN, int(T * cfg.MEM_MAX_LEN_RATIO) + cfg.DEC_MAX_LEN_PAD)
        return max(a, b, 1)

    def decode_greedy(self, mem_bf16: torch.Tensor, len_est: torch.Tensor, B: int, T: int, Lmax: int,
                      select_raw: bool = False, forced: Optional[torch.Tensor] = None, want_steps: bool = False,
                      poll_every: int = 8):
        p = self.decode_params(select_raw)
        need = self.lib.kiri_decode_workspace_bytes(self.handle, B, T, Lmax)
        ws = self._workspace(need, "_dws")
        ids = torch.zeros((B, Lmax), dtype=torch.int32, device=self.device)
        n_out = torch.zeros(B, dtype=torch.int32, device=self.device)
        sum_lp = torch.zeros(B, dtype=torch.float32, device=self.device)
        slp = torch.zeros((B, Lmax), dtype=torch.float32, device=self.device) if want_steps else None
        spr = torch.zeros((B, Lmax), dtype=torch.float32, device=self.device) if want_steps else None
        steps = C.c_int(0)
        _lib.check(self.lib.kiri_decode_greedy(self.handle, mem_bf16.data_ptr(), len_est.data_ptr(), B, T, Lmax,
                                               C.byref(p), ws.data_ptr(), need, ids.data_ptr(), n_out.data_ptr(),
                                               sum_lp.data_ptr(), _lib.ptr(slp), _lib.ptr(spr), _lib.ptr(forced),
                                               C.byref(steps), poll_every, _lib.stream_ptr()), "kiri_decode_greedy")
        self.launches += 3          # cross-K/V GEMM, head-major relayout, the persistent decode kernel
        return ids, n_out, sum_lp, slp, spr, steps.value

    # ------------------------------------------------------------------ device-resident stepping (bench)
    def prepare_resident(self, src: torch.Tensor, entries: np.ndarray):
        """Upload the source buffer and the crop descriptors once; returns an opaque plan that
        ``step_resident`` replays with no host<->device traffic."""
        src_dev = src if src.is_cuda else src.to(self.device)
        plan = []
        row0, p0 = 0, 0
        mem_row0, mem_len, kv, dall = [], [], [], []
        groups = self.plan(entries)
        total_planes = sum(len(g[0]) * self.cfg.IMG_H * Wb for Wb, g in groups.items())
        planes_all = torch.empty(total_planes, dtype=torch.uint8, device=self.device)
        smem_max = n_strips_max = 0
        for Wb, (idx, descs, smem, n_strips) in groups.items():
            descs = descs.copy()
            descs["out_offset"] += p0                                   # absolute inside planes_all
            dall.append(descs)
            nb = len(idx) * self.cfg.IMG_H * Wb
            planes = planes_all[p0:p0 + nb].view(len(idx), self.cfg.IMG_H, Wb)
            p0 += nb
            smem_max, n_strips_max = max(smem_max, smem), max(n_strips_max, n_strips)
            T = Wb // 4
            mem_row0.append(row0 + np.arange(len(idx), dtype=np.int32) * T)
            mem_len.append(np.full(len(idx), T, np.int32))
            kv.append(np.minimum((descs["nw"] + 3) // 4, T).astype(np.int32))
            row0 += len(idx) * T
            plan.append({"Wb": Wb, "idx": idx, "n": len(idx), "planes": planes})
        dd = torch.from_numpy(np.concatenate(dall).view(np.uint8).reshape(-1).copy()).to(self.device)
        kv_len = torch.from_numpy(np.concatenate(kv)).to(self.device) if self.width_mode == "masked" else None
        out = {"src": src_dev, "groups": plan, "kv_len": kv_len, "descs": dd,
               "sums": torch.empty(2 * len(entries), dtype=torch.int32, device=self.device), "n_crops": len(entries), "smem": smem_max,
               "n_strips": n_strips_max, "planes_all": planes_all, "M": row0, "n_lines": len(entries),
               "mem_row0": torch.from_numpy(np.concatenate(mem_row0)).to(self.device),
               "mem_len": torch.from_numpy(np.concatenate(mem_len)).to(self.device),
               "T_max": max(g["Wb"] for g in plan) // 4}
        torch.cuda.synchronize()
        return out

    def step_resident(self, prep, method: str = "ctc"):
        """One pass of the hot path over the prepared batch, inputs already in HBM.  Returns the
        device outputs; nothing synchronises with the host ("ctc" and "decoder")."""
        _lib.check(self.lib.kiri_preprocess_pack(prep["src"].data_ptr(), prep["descs"].data_ptr(), prep["n_crops"],
                                                 self.cfg.IMG_H, prep["smem"], prep["n_strips"],
                                                 prep["planes_all"].data_ptr(), 0, prep["sums"].data_ptr(), _lib.stream_ptr()),
                   "kiri_preprocess_pack")
        self.launches += 2
        M, L = prep["M"], prep["n_lines"]
        fid = torch.empty(M, dtype=torch.int32, device=self.device)
        fpr = torch.empty(M, dtype=torch.float32, device=self.device)
        # frame decisions come out of the CTC head's epilogue: no logits in HBM for "ctc" / "decoder"
        enc = self.encode_multi([g["planes"] for g in prep["groups"]], kv_len=prep["kv_len"], want_logits=False, stats=(fid, fpr),
                                want_mem=(method == "decoder"))
        ids = torch.empty(M, dtype=torch.int32, device=self.device)
        n_ids = torch.empty(L, dtype=torch.int32, device=self.device)
        conf = torch.empty(L, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.kiri_ctc_collapse_multi(fid.data_ptr(), fpr.data_ptr(), L, prep["mem_row0"].data_ptr(),
                                                    prep["mem_len"].data_ptr(), ids.data_ptr(), n_ids.data_ptr(), conf.data_ptr(),
                                                    _lib.stream_ptr()), "kiri_ctc_collapse_multi")
        self.launches += 1
        outs = [(ids, n_ids, conf)]
        if method != "decoder":
            return outs
        # the decode loop's capacity is a static bound and every line's own step budget is derived on the device from
        # its length estimate: no host synchronisation between the encoder and the decoder
        Lmax = self.static_step_cap(prep["T_max"])
        d_ids, n_out, sum_lp, _, _ = self.decode_greedy_multi(enc["mem_bf16"], prep["mem_row0"], prep["mem_len"], n_ids, Lmax,
                                                              prep["T_max"])
        return [(d_ids, n_out, sum_lp, conf)]

    def profile(self, fn):
        """Run ``fn()`` with per-stage CUDA-event timing on; returns {stage: (ms, intervals)}."""
        n = self.lib.kiri_profile_begin()
        try:
            fn()
        finally:
            ms = (C.c_double * n)()
            cnt = (C.c_int * n)()
            _lib.check(self.lib.kiri_profile_end(ms, cnt, n), "kiri_profile_end")
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROFILE_STAGES[:n])}

    # ------------------------------------------------------------------ public API
    def _check_device(self):
        if torch.cuda.current_device() != self.device.index:
            raise _lib.KiriError(f"this engine lives on {self.device}; the current device is cuda:{torch.cuda.current_device()} "
                                 f"(wrap the call in torch.cuda.device({self.device.index}))")

    def static_step_cap(self, T_max: int) -> int:
        """Capacity of the decode loop that needs no host round trip: the CTC length estimate is at most T_max
        (one new id per frame), so max_steps (model.py:416-425) is at most this.  The kernel derives every line's own
        max_steps from its device-resident length estimate and clusters exit as soon as their lines are done."""
        return self.max_steps_bound(T_max, T_max)

    def _stage_host(self, arrays: Sequence[np.ndarray], slot: int, align: int = 1):
        """Concatenate host arrays into the slot's PERSISTENT pinned buffer (no per-call cudaHostAlloc); with
        ``align`` > 1 every array starts at a multiple of it.  Returns (buffer view, start offset of every array)."""
        offs, total = [], 0
        for a in arrays:
            total = -(-total // align) * align
            offs.append(total)
            total += int(a.size)
        name = f"_hsrc_{slot}"
        cur = getattr(self, name, None)
        if cur is None or cur.numel() < total + 16:
            cur = torch.empty(int((total + 16) * 1.25) + 4096, dtype=torch.uint8).pin_memory()
            setattr(self, name, cur)
        elif self._h2d_done[slot] is not None:
            self._h2d_done[slot].synchronize()              # the upload that last read this buffer has finished
        nb = cur.numpy()
        for a, off in zip(arrays, offs):
            if a.dtype != np.uint8:
                raise ValueError("source arrays must be uint8")
            nb[off:off + a.size] = a.reshape(-1)
        return cur[:total + 16], offs

    @torch.no_grad()
    def submit(self, src, entries: np.ndarray, method: str = "ctc", streaming: bool = False, bgr_pages=None,
               live: bool = False):
        """Enqueue one batch (see ``_submit_own``) on the engine's stream; the caller's stream is waited for first."""
        self._check_device()
        caller = torch.cuda.current_stream(self.device)
        if caller != self.stream:
            self.stream.wait_stream(caller)
            with torch.cuda.stream(self.stream), _lib.pinned_stream(self.stream.cuda_stream):
                return self._submit_own(src, entries, method, streaming, bgr_pages, live)
        with _lib.pinned_stream(self.stream.cuda_stream):
            return self._submit_own(src, entries, method, streaming, bgr_pages, live)

    def _submit_own(self, src, entries: np.ndarray, method: str = "ctc", streaming: bool = False, bgr_pages=None,
               live: bool = False):
        """Enqueue one batch (H2D of the source on the copy stream, preprocess, encoder, CTC greedy, for "decoder"
        the whole greedy decode, async D2H of the packed results) and return a ticket without synchronising the
        host.  Two tickets may be in flight: ``t2 = submit(...); r1 = collect(t1)`` overlaps batch i's host-side
        string decoding and batch i+1's upload with the kernels of the other batch.
        ``src``: uint8 buffer (pinned host or device) holding pages/crops, or a list of uint8 numpy arrays that are
        concatenated into the engine's persistent pinned staging; ``entries[n,4]`` = (byte offset, pitch, w, h)
        of every crop after the reference's clamp-pad; an optional fifth column carries ``KiriCropDesc.flags``
        (``_lib.CROP_NO_INVERT`` for inputs that are already preprocessed planes)."""
        if method not in ("ctc", "decoder", "beam"):
            raise ValueError("method must be 'ctc', 'decoder' or 'beam'")
        self._check_device()
        n = len(entries)
        live = live and method in ("decoder", "beam")
        tk = {"method": method, "streaming": streaming, "n": n, "live": live}
        if n == 0:
            return tk
        marks = [_time.perf_counter()]                       # host-side phase marks of this call (tk["marks"])
        self._slot ^= 1
        slot, sl = self._slot, f"_{self._slot}"
        if isinstance(src, (list, tuple)):
            src, _ = self._stage_host(src, slot)
        if src.is_cuda:
            src_dev = src
        else:
            # upload on the copy stream: it overlaps the kernels of the batch submitted before
            had = getattr(self, "_src" + sl, None)
            src_dev = self._device("_src" + sl, src.numel(), torch.uint8)
            if had is not src_dev:
                # a fresh block from the caching allocator may have been freed on the engine stream a moment ago
                # (e.g. a workspace that grew) while kernels that read it are still queued there
                self._copy_stream.wait_stream(self.stream)
            if self._src_free[slot] is not None:              # the preprocess launch that last read this slot
                self._copy_stream.wait_event(self._src_free[slot])
            with torch.cuda.stream(self._copy_stream):
                src_dev[:src.numel()].copy_(src, non_blocking=True)
                self._h2d_done[slot] = torch.cuda.Event()
                self._h2d_done[slot].record()
        if bgr_pages:
            # GPU-side page ingest: `src` holds raw pages, some of them interleaved BGR; the
            # gray pages the crops are cut from are produced on the device (cv2-exact kernel), entries address them
            total_gray = max(g + n_px for _, g, n_px, _ in bgr_pages)
            gsrc = self._device("_gsrc" + sl, total_gray + 16, torch.uint8)
            if not src.is_cuda:
                self.stream.wait_stream(self._copy_stream)
            for raw_off, gray_off, n_px, is_bgr in bgr_pages:
                if is_bgr:
                    _lib.check(self.lib.kiri_bgr_to_gray(src_dev.data_ptr() + raw_off, n_px, gsrc.data_ptr() + gray_off,
                                                         _lib.stream_ptr()), "kiri_bgr_to_gray")
                    self.launches += 1
                else:
                    gsrc[gray_off:gray_off + n_px].copy_(src_dev[raw_off:raw_off + n_px], non_blocking=True)
            src_dev = gsrc
        marks.append(_time.perf_counter())                   # [1] upload enqueued
        groups = list(self.plan(entries).items())
        marks.append(_time.perf_counter())                   # [2] planned
        IMG_H, Cp = self.cfg.IMG_H, self.pw.Cp
        n_lines = sum(len(g[1][0]) for g in groups)
        M = sum(len(g[0]) * (Wb // 4) for Wb, g in groups)
        # ---- ONE pinned staging buffer -> ONE H2D copy for every descriptor / per-line table
        desc_bytes = sum(g[1][1].nbytes for g in groups)
        meta_words = desc_bytes // 4 + 3 * n_lines                       # descs | kv_len | mem_row0 | mem_len
        hmeta = self._pinned("_hmeta" + sl, meta_words, torch.int32)
        hm = hmeta.numpy()
        off, line0, row0, p0 = 0, 0, 0, 0
        kvo, r0o, mlo = desc_bytes // 4, desc_bytes // 4 + n_lines, desc_bytes // 4 + 2 * n_lines
        planes_list = []
        planes_all = self._device("_planes" + sl, M * 4 * IMG_H, torch.uint8)
        smem_max = n_strips_max = 0
        for Wb, (idx, descs, smem, n_strips) in groups:
            nb, T = descs.nbytes // 4, Wb // 4
            descs["out_offset"] += p0                                   # absolute inside planes_all
            hm[off:off + nb] = descs.view(np.int32).reshape(-1)
            hm[kvo + line0:kvo + line0 + len(idx)] = np.minimum((descs["nw"] + 3) // 4, T)
            hm[r0o + line0:r0o + line0 + len(idx)] = row0 + np.arange(len(idx), dtype=np.int32) * T
            hm[mlo + line0:mlo + line0 + len(idx)] = T
            npl = len(idx) * IMG_H * Wb
            planes_list.append(planes_all[p0:p0 + npl].view(len(idx), IMG_H, Wb))
            p0 += npl
            smem_max, n_strips_max = max(smem_max, smem), max(n_strips_max, n_strips)
            off += nb
            line0 += len(idx)
            row0 += len(idx) * T
        dmeta = self._device("_dmeta" + sl, meta_words, torch.int32)
        dmeta[:meta_words].copy_(hmeta[:meta_words], non_blocking=True)
        if not src.is_cuda:
            self.stream.wait_stream(self._copy_stream)
        marks.append(_time.perf_counter())                   # [3] staging filled, descriptors enqueued
        # ---- ONE preprocess launch for every width group, one encoder pass over all groups
        sums = self._device("_sums" + sl, 2 * n_lines, torch.int32)
        _lib.check(self.lib.kiri_preprocess_pack(src_dev.data_ptr(), dmeta.data_ptr(), n_lines, IMG_H, smem_max, n_strips_max,
                                                 planes_all.data_ptr(), 0, sums.data_ptr(), _lib.stream_ptr()),
                   "kiri_preprocess_pack")
        self.launches += 2
        if not src.is_cuda:
            self._src_free[slot] = torch.cuda.Event()
            self._src_free[slot].record()
        kv_len = dmeta[kvo:kvo + n_lines] if self.width_mode == "masked" else None
        marks.append(_time.perf_counter())                   # [4] preprocess launched
        # ---- ONE packed result buffer: ids[M] | n_ids[L] | conf[L] | decoder block | frame ids[M] | frame probs[M]
        # (the frame decisions are taken in the CTC head's GEMM epilogue; they are downloaded only for CTC streaming)
        want_frames = streaming and method == "ctc"
        T_max = max(Wb // 4 for Wb, _ in groups)
        Lcap = self.static_step_cap(T_max) if (method == "decoder" or live) else 0
        LL = n_lines * Lcap
        dec_words = (3 * LL if streaming else 2 * LL) + 2 * n_lines if (method == "decoder" and not live) else 0
        ctc_words = M + 2 * n_lines
        head_words = ctc_words + dec_words
        res_words = head_words + (2 * M if want_frames else 0)
        dres = self._device("_dres" + sl, head_words + 2 * M, torch.int32)
        ids_all, n_all = dres[:M], dres[M:M + n_lines]
        conf_all = dres[M + n_lines:M + 2 * n_lines].view(torch.float32)
        fid_all = dres[head_words:head_words + M]
        fpr_all = dres[head_words + M:head_words + 2 * M].view(torch.float32)
        enc = self.encode_multi(planes_list, kv_len=kv_len, slot=sl, want_logits=(method == "beam"), stats=(fid_all, fpr_all),
                                want_mem=(method != "ctc" or live))
        marks.append(_time.perf_counter())                   # [5] encoder launched
        _lib.check(self.lib.kiri_ctc_collapse_multi(fid_all.data_ptr(), fpr_all.data_ptr(), n_lines, dmeta[r0o:].data_ptr(),
                                                    dmeta[mlo:].data_ptr(), ids_all.data_ptr(), n_all.data_ptr(),
                                                    conf_all.data_ptr(), _lib.stream_ptr()), "kiri_ctc_collapse_multi")
        self.launches += 1
        mem_row0, mem_len = dmeta[r0o:r0o + n_lines], dmeta[mlo:mlo + n_lines]
        if live:
            tk["live_out"] = self._launch_live(tk, enc, mem_row0, mem_len, n_all, n_lines, Lcap, T_max, method, sl)
        elif method == "decoder":
            # the greedy decode is enqueued right behind the CTC kernel: its step bounds come from the device-resident
            # length estimates, so no host round trip separates the encoder from the decoder
            dd = dres[ctc_words:head_words]
            d_ids, n_out = dd[:LL].view(n_lines, Lcap), dd[LL:LL + n_lines]
            sum_lp = dd[LL + n_lines:LL + 2 * n_lines].view(torch.float32)
            slp = dd[LL + 2 * n_lines:2 * LL + 2 * n_lines].view(torch.float32).view(n_lines, Lcap)
            spr = dd[2 * LL + 2 * n_lines:3 * LL + 2 * n_lines].view(torch.float32).view(n_lines, Lcap) if streaming else None
            self.decode_greedy_multi(enc["mem_bf16"], mem_row0, mem_len, n_all, Lcap, T_max, select_raw=streaming,
                                     out=(d_ids, n_out, sum_lp, slp, spr))
        hres = self._pinned("_hres" + sl, res_words, torch.int32)
        hres[:res_words].copy_(dres[:res_words], non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        marks.append(_time.perf_counter())                   # [6] CTC (+ decode) + download enqueued
        tk.update(marks=marks, done=done, hres=hres, res_words=res_words, ctc_words=ctc_words, head_words=head_words, Lcap=Lcap, M=M,
                  n_lines=n_lines,
                  rows=enc["rows"], T_max=T_max,
                  order=np.concatenate([g[1][0] for g in groups]),     # line index of every concatenated slot
                  want_frames=want_frames, enc=enc, n_all=n_all, mem_row0=mem_row0, mem_len=mem_len,
                  keep=(src_dev, planes_all, dmeta, dres), sl=sl)
        return tk

    @torch.no_grad()
    def collect(self, tk) -> List[Optional[LineResult]]:
        """Wait for a ticket and build the per-line results on the host ("beam" runs its search now: its Python-side
        final ranking needs the CTC length estimates for the buffer sizes)."""
        n = tk["n"]
        results: List[Optional[LineResult]] = [None] * n
        if n == 0:
            return results
        self._check_device()
        caller = torch.cuda.current_stream(self.device)
        if caller != self.stream:
            with torch.cuda.stream(self.stream):
                out = self.collect(tk)
            caller.wait_stream(self.stream)
            return out
        with _lib.pinned_stream(self.stream.cuda_stream):
            return self._collect_own(tk, results)

    def _collect_own(self, tk, results):
        n = tk["n"]
        method, streaming = tk["method"], tk["streaming"]
        tk["done"].synchronize()
        res_words, M, n_lines, order, enc = tk["res_words"], tk["M"], tk["n_lines"], tk["order"], tk["enc"]
        hr = tk["hres"].numpy()[:res_words].copy()                  # the pinned buffer is reused two batches later
        n_h = hr[M:M + n_lines]
        c_h = hr[M + n_lines:M + 2 * n_lines].view(np.float32)
        if method == "ctc":
            want_frames = tk["want_frames"]
            # texts of the whole batch in one vectorised pass (CharTokenizer.decode_batch)
            flat = [hr[r0:r0 + B * T].reshape(B, T)[np.arange(T)[None, :] < n_h[p0:p0 + B, None]]
                    for (r0, B, T), p0 in zip(tk["rows"], np.cumsum([0] + [b for _, b, _ in tk["rows"]])[:-1])]
            texts = self.tok.decode_batch(np.concatenate(flat), n_h, "ctc")
            pos = 0
            n_l, c_l, o_l = n_h.tolist(), c_h.tolist(), (order.tolist() if hasattr(order, "tolist") else list(order))
            hw = tk["head_words"]
            for (r0, B, T) in tk["rows"]:
                ids_h = hr[r0:r0 + B * T].reshape(B, T)
                if want_frames:
                    f_h = hr[hw + r0:hw + r0 + B * T].reshape(B, T)
                    p_h = hr[hw + M + r0:hw + M + r0 + B * T].view(np.float32).reshape(B, T)
                    for j in range(B):
                        k = pos + j
                        results[o_l[k]] = LineResult(texts[k], c_l[k], c_l[k], ids_h[j, :n_l[k]], None, None, f_h[j], p_h[j], n_l[k])
                else:                                                # (python scalars and positional arguments: this loop is
                    for j in range(B):                               #  a third of collect()'s host time at 256 lines)
                        k = pos + j
                        results[o_l[k]] = LineResult(texts[k], c_l[k], c_l[k], ids_h[j, :n_l[k]], None, None, None, None, n_l[k])
                pos += B
            return results
        len_h = n_h                                                  # length estimates bound the loop
        if method == "beam":
            return self._beam_finish(enc, tk["mem_row0"], tk["mem_len"], tk["n_all"], len_h, c_h, order, results)
        # ---- greedy attention decoder: everything is already on the host
        Lcap, LL = tk["Lcap"], n_lines * tk["Lcap"]
        hd = hr[tk["ctc_words"]:tk["head_words"]]
        ids_h, no_h = hd[:LL].reshape(n_lines, Lcap), hd[LL:LL + n_lines]
        slp_h = hd[LL + 2 * n_lines:2 * LL + 2 * n_lines].view(np.float32).reshape(n_lines, Lcap)
        spr_h = hd[2 * LL + 2 * n_lines:3 * LL + 2 * n_lines].view(np.float32).reshape(n_lines, Lcap) if streaming else None
        eos = self.tok.dec_eos
        col = np.arange(Lcap)[None, :]
        valid = col < no_h[:, None]
        is_eos = (ids_h == eos) & valid
        cut = np.where(is_eos.any(1), is_eos.argmax(1), no_h)            # ids before the first EOS make the text
        texts = self.tok.decode_batch(ids_h[col < cut[:, None]], cut, "dec")
        lp_sum = np.where(valid, slp_h, 0.0).astype(np.float64).sum(1)
        for j, li in enumerate(order):
            nj = int(no_h[j])
            dec_conf = min(1.0, max(0.0, math.exp(float(lp_sum[j]) / nj))) if nj else 0.0
            results[li] = LineResult(texts[j], 0.6 * dec_conf + 0.4 * float(c_h[j]), float(c_h[j]), ids_h[j, :nj],
                                     step_logp=slp_h[j, :nj], step_prob=None if spr_h is None else spr_h[j, :nj],
                                     len_est=int(len_h[j]))
        return results

    # ------------------------------------------------------------------ live streaming
    def _launch_live(self, tk, enc, mem_row0, mem_len, len_est, n_lines, Lcap, T_max, method, sl):
        """Enqueue the decode with its outputs in MAPPED PINNED HOST memory and per-step publication: the host reads
        tokens while the persistent kernel is still decoding (later regions decode ahead of the one being read)."""
        LL = n_lines * Lcap
        if method == "decoder":
            words = 3 * LL + 3 * n_lines                     # ids | step_logp | step_prob | n_out | sum_lp | progress
        else:
            beam = int(self.cfg.BEAM)
            if not 1 <= beam <= 5:
                raise ValueError(f"cfg.BEAM={beam}: the B200 beam decoder supports widths 1..5")
            words = 3 * LL * beam + n_lines                  # trace [n, Lcap, beam, 3] | progress
        h = self._pinned("_hlive" + sl, words, torch.int32)
        hv = h.numpy()
        base = h.data_ptr()                                  # UVA: the device writes through the same address
        if method == "decoder":
            po = 3 * LL + 2 * n_lines
            hv[po:po + n_lines] = 0
            out = (_Raw(base), _Raw(base + 4 * (3 * LL)), _Raw(base + 4 * (3 * LL + n_lines)), _Raw(base + 4 * LL), _Raw(base + 8 * LL))
            self.decode_greedy_multi(enc["mem_bf16"], mem_row0, mem_len, len_est, Lcap, T_max, select_raw=True, out=out,
                                     progress_ptr=base + 4 * po, publish=True)
            return {"ids": hv[:LL].reshape(n_lines, Lcap), "logp": hv[LL:2 * LL].view(np.float32).reshape(n_lines, Lcap),
                    "prob": hv[2 * LL:3 * LL].view(np.float32).reshape(n_lines, Lcap), "progress": hv[po:po + n_lines]}
        po = 3 * LL * beam
        hv[po:po + n_lines] = 0
        M = int(enc["mem_bf16"].shape[0])
        p = self.decode_params(False)
        need = self.lib.kiri_decode_beam_workspace_bytes(self.handle, n_lines, M, Lcap, beam)
        ws = self._workspace(need, "_dws")
        nb = n_lines * beam
        dout = self._device("_dbeam", 5 * nb + 2 * nb * Lcap, torch.int32)
        dout[:5 * nb].zero_()
        perm = torch.argsort(len_est, descending=True, stable=True).to(torch.int32)
        _lib.check(self.lib.kiri_decode_beam_multi(
            self.handle, enc["mem_bf16"].data_ptr(), M, mem_row0.data_ptr(), mem_len.data_ptr(), T_max, len_est.data_ptr(),
            perm.data_ptr(), n_lines, Lcap, beam, float(self.cfg.BEAM_LENP), C.byref(p), ws.data_ptr(), need,
            dout[:2 * nb].data_ptr(), dout[2 * nb:].data_ptr(), dout[3 * nb:].data_ptr(), dout[5 * nb:].data_ptr(),
            dout[5 * nb + nb * Lcap:].data_ptr(), 1, base, base + 4 * po, 1, _lib.stream_ptr()), "kiri_decode_beam_multi")
        self.launches += 2
        return {"trace": hv[:po].reshape(n_lines, Lcap, beam, 3), "progress": hv[po:po + n_lines], "beam": beam}

    @staticmethod
    def _poll(progress, k, seen, timeout_s=120.0):
        """Wait until slot k has more than `seen` steps or has ended; returns (steps available, ended)."""
        t0 = _time.perf_counter()
        while True:
            v = int(progress[k])
            n, done = v & 0x3FFFFFFF, bool(v >> 30)
            if n > seen or done:
                return n, done
            if _time.perf_counter() - t0 > timeout_s:
                raise _lib.KiriError("live decode: no progress from the device (kernel fault?)")
            _time.sleep(0)

    def live_slot(self, tk, line: int) -> int:
        """Decode slot of line `line` of a live ticket (results are stored in width-group order)."""
        inv = tk.get("_inv")
        if inv is None:
            inv = np.empty(tk["n"], np.int64)
            inv[tk["order"]] = np.arange(tk["n"])
            tk["_inv"] = inv
        return int(inv[line])

    def live_greedy(self, tk, line: int):
        """Generator over the tokens of one line AS THE DEVICE PRODUCES THEM: (token id, raw soft-max probability,
        penalised log-prob, step, finished).  The streaming token rule applies (arg-max of the raw dec_head soft-max,
        model.py:915-917)."""
        lo, k, seen = tk["live_out"], self.live_slot(tk, line), 0
        eos = self.tok.dec_eos
        while True:
            n, ended = self._poll(lo["progress"], k, seen)
            for s in range(seen, n):
                tid = int(lo["ids"][k, s])
                yield tid, float(lo["prob"][k, s]), float(lo["logp"][k, s]), s + 1, tid == eos
            seen = n
            if ended:
                return

    def live_beam(self, tk, line: int):
        """Generator over the decode steps of one line in beam-streaming mode (model.py:949-1152): after every step the
        BEST partial hypothesis as (ids after BOS, their log-probs, step, finished), rebuilt from the device's trace."""
        lo, k, seen = tk["live_out"], self.live_slot(tk, line), 0
        beam, eos = lo["beam"], self.tok.dec_eos
        hyps = [([], [])] + [None] * (beam - 1)
        while True:
            n, ended = self._poll(lo["progress"], k, seen)
            for s in range(seen, n):
                tr = lo["trace"][k, s]
                new = [None] * beam
                for r in range(beam):
                    par, tokid, bits = int(tr[r, 0]), int(tr[r, 1]), tr[r, 2:3]
                    if par < 0 or hyps[par] is None:
                        continue
                    ids, lps = hyps[par]
                    new[r] = (ids + [tokid], lps + [float(bits.view(np.float32)[0])]) if tokid >= 0 else (ids, lps)
                hyps = new
                ids, lps = hyps[0]
                yield ids, lps, s + 1, bool(ids) and ids[-1] == eos
            seen = n
            if ended:
                return

    def live_finish(self, tk):
        """Wait for a live ticket's kernels (call when the consumer stops reading, so buffers can be reused)."""
        if tk.get("n"):
            tk["done"].synchronize()

    def ticket_records(self, tk, T: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The exchange records of a submitted batch, built ON THE DEVICE from the ticket's outputs (stream-ordered
        behind its kernels, no host involvement): int32 [n_lines, 2 + T] = {n_ids, confidence bits, ids[T]} in the
        batch's slot order ("ctc": collapsed CTC ids + CTC confidence; "decoder": decoder ids + sum of log-probs)."""
        n, M = tk["n_lines"], tk["M"]
        rec = out if out is not None else torch.zeros((n, 2 + T), dtype=torch.int32, device=self.device)
        dres = tk["keep"][3]
        if tk["method"] == "decoder":
            Lcap, LL = tk["Lcap"], n * tk["Lcap"]
            dd = dres[tk["ctc_words"]:tk["head_words"]]
            k = min(T, Lcap)
            rec[:n, 0] = dd[LL:LL + n]
            rec[:n, 1] = dd[LL + n:LL + 2 * n]
            rec[:n, 2:2 + k] = dd[:LL].view(n, Lcap)[:, :k]
            self.launches += 3
        else:
            _lib.check(self.lib.kiri_pack_records(dres[:M].data_ptr(), dres[M:M + n].data_ptr(), dres[M + n:M + 2 * n].data_ptr(),
                                                  tk["mem_row0"].data_ptr(), n, T, rec.data_ptr(), _lib.stream_ptr()),
                       "kiri_pack_records")
            self.launches += 1
        return rec

    def recognize_packed(self, src: torch.Tensor, entries: np.ndarray, method: str = "ctc",
                         streaming: bool = False) -> List[Optional[LineResult]]:
        """One synchronous batch = ``collect(submit(...))``."""
        return self.collect(self.submit(src, entries, method, streaming))

    def _beam_finish(self, enc, mem_row0, mem_len, len_est, len_h, ctc_conf_h, order, results):
        """decode_method="beam" (model.py:390-600 with cfg.BEAM > 1): device beam search + device CTC
        forward scores, then the reference's final ranking (model.py:562-598) in Python floats."""
        cfg, tok = self.cfg, self.tok
        beam = int(cfg.BEAM)
        if not 1 <= beam <= 5:
            raise ValueError(f"cfg.BEAM={beam}: the B200 beam decoder supports widths 1..5")
        n_lines, M = len(order), int(enc["mem_bf16"].shape[0])
        T_max = max(T for _, _, T in enc["rows"])
        Lmax = self.max_steps_bound(int(len_h.max()), T_max)
        p = self.decode_params(False)
        need = self.lib.kiri_decode_beam_workspace_bytes(self.handle, n_lines, M, Lmax, beam)
        ws = self._workspace(need, "_dws")
        nb = n_lines * beam
        # packed outputs: score f64[nb] | len i32[nb] | state i32[nb] | align f32[nb] | ids i32[nb*Lmax] | logp f32[nb*Lmax]
        words = 2 * nb + 3 * nb + 2 * nb * Lmax
        dout = self._device("_dbeam", words, torch.int32)
        dout[:5 * nb].zero_()
        score = dout[:2 * nb].view(torch.float64)
        blen, bstate = dout[2 * nb:3 * nb], dout[3 * nb:4 * nb]
        align = dout[4 * nb:5 * nb].view(torch.float32)
        bids = dout[5 * nb:5 * nb + nb * Lmax]
        blp = dout[5 * nb + nb * Lmax:5 * nb + 2 * nb * Lmax].view(torch.float32)
        perm = torch.argsort(len_est, descending=True, stable=True).to(torch.int32)
        _lib.check(self.lib.kiri_decode_beam_multi(self.handle, enc["mem_bf16"].data_ptr(), M, mem_row0.data_ptr(),
                                                   mem_len.data_ptr(), T_max, len_est.data_ptr(), perm.data_ptr(), n_lines, Lmax,
                                                   beam, float(cfg.BEAM_LENP), C.byref(p), ws.data_ptr(), need,
                                                   score.data_ptr(), blen.data_ptr(), bstate.data_ptr(), bids.data_ptr(),
                                                   blp.data_ptr(), 0, 0, 0, 0, _lib.stream_ptr()), "kiri_decode_beam_multi")
        fuse_ctc = cfg.USE_CTC and cfg.CTC_FUSION_ALPHA > 0
        if fuse_ctc:
            _lib.check(self.lib.kiri_ctc_align_score(enc["logits"].data_ptr(), self.pw.Cp, self.pw.C, mem_row0.data_ptr(),
                                                     mem_len.data_ptr(), n_lines, beam, Lmax, bids.data_ptr(),
                                                     blen.data_ptr(), bstate.data_ptr(), tok.vocab_size,
                                                     tok.unk_id + tok.ctc_offset, T_max, align.data_ptr(),
                                                     _lib.stream_ptr()), "kiri_ctc_align_score")
        self.launches += 3
        hout = self._pinned("_hbeam", words, torch.int32)
        hout[:words].copy_(dout[:words], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = hout.numpy()[:words].copy()
        sc_h = h[:2 * nb].view(np.float64).reshape(n_lines, beam)
        ln_h = h[2 * nb:3 * nb].reshape(n_lines, beam)
        st_h = h[3 * nb:4 * nb].reshape(n_lines, beam)
        al_h = h[4 * nb:5 * nb].view(np.float32).reshape(n_lines, beam)
        id_h = h[5 * nb:5 * nb + nb * Lmax].reshape(n_lines, beam, Lmax)
        lp_h = h[5 * nb + nb * Lmax:].view(np.float32).reshape(n_lines, beam, Lmax)
        tab, eos = self._dec_table, tok.dec_eos
        for j, li in enumerate(order):
            best, best_key = None, None
            for r in range(beam):                                   # hypotheses are in pruned (normed) order
                if st_h[j, r] == 0:
                    continue
                n = int(ln_h[j, r])
                length = max(1, n)
                dec_score = float(sc_h[j, r]) / (length ** cfg.BEAM_LENP if length > 0 else 1.0)
                key = dec_score + cfg.CTC_FUSION_ALPHA * float(al_h[j, r]) if fuse_ctc else dec_score
                if best is None or key > best_key:                  # stable: ties keep the earlier hypothesis
                    best, best_key = r, key
            n = int(ln_h[j, best])
            row = id_h[j, best, :n]
            lps = lp_h[j, best, :n].astype(np.float64)
            dec_conf = min(1.0, max(0.0, math.exp(float(lps.sum()) / n))) if n else 0.0
            hit = np.nonzero(row == eos)[0]
            text_ids = row[:hit[0]] if len(hit) else row
            cc = float(ctc_conf_h[j])
            results[li] = LineResult("".join([tab[i] for i in text_ids.tolist()]), 0.6 * dec_conf + 0.4 * cc, cc, row,
                                     step_logp=lp_h[j, best, :n], step_prob=np.exp(lp_h[j, best, :n]), len_est=int(len_h[j]))
        return results

    def recognize_crops(self, crops: Sequence[np.ndarray], method: str = "ctc", streaming: bool = False):
        """Line crops (2-D uint8 arrays) -> results, through the engine's persistent pinned staging."""
        ent = np.zeros((len(crops), 4), np.int64)
        off = 0
        arrs = []
        for i, c in enumerate(crops):
            if c.dtype != np.uint8 or c.ndim != 2:
                raise ValueError("crops must be 2-D uint8 arrays")
            ent[i] = (off, c.shape[1], c.shape[1], c.shape[0])
            off += c.size
            arrs.append(np.ascontiguousarray(c))
        return self.collect(self.submit(arrs, ent, method, streaming))

    @staticmethod
    def _page_entries(page_shape, boxes, page_offset: int = 0):
        """Clamp-padded entries of one page's boxes.  A box that cannot be interpreted (wrong arity, non-numeric)
        becomes a LineError instead of failing the page (per-region isolation, core.py:771-791); returns
        (entries of the usable boxes, their indices, {index: LineError})."""
        errors = {}
        good, rows = [], []
        for i, b in enumerate(boxes):
            try:
                x, y, w, h = (int(v) for v in b)
                rows.append((x, y, w, h))
                good.append(i)
            except Exception as e:                              # noqa: BLE001 - mirrored from the reference's bare except
                errors[i] = LineError(f"invalid box {b!r}: {e}")
        if not rows:
            return np.zeros((0, 4), np.int64), np.zeros(0, np.int64), errors
        ent, valid = BatchedRecognizer.boxes_to_entries(page_shape, rows, page_offset)
        return ent[valid], np.asarray(good, np.int64)[valid], errors

    def _isolate(self, arrays, ent, method, streaming, exc, bgr=None):
        """The batch failed as a whole: recognise its lines one by one so that a single bad region cannot take the
        others down (the reference's per-region try/except); a line that fails alone becomes a LineError."""
        out = []
        for k in range(len(ent)):
            try:
                out.append(self.collect(self.submit(arrays, ent[k:k + 1], method, streaming, bgr))[0])
            except Exception as e:                              # noqa: BLE001
                out.append(LineError(str(e)))
        if all(isinstance(r, LineError) for r in out):
            raise exc                                           # nothing works: a device/library problem, not a bad line
        return out

    def recognize_boxes(self, page_gray: np.ndarray, boxes: Sequence[Sequence[int]], method: str = "ctc",
                        streaming: bool = False) -> List[Optional[LineResult]]:
        """All boxes of one grayscale page; boxes whose clamped crop is empty give ``None``
        (the reference skips them, core.py:516-517 / 773-774), malformed boxes a ``LineError``."""
        return self.recognize_pages([page_gray], [boxes], method, streaming)[0]

    def recognize_pages(self, pages: Sequence[np.ndarray], boxes_list: Sequence[Sequence[Sequence[int]]],
                        method: str = "ctc", streaming: bool = False, batch_lines: int = 384):
        """Several pages with their detector boxes: whole pages are grouped into batches of about ``batch_lines``
        lines, every page is uploaded ONCE (crops are taken on the device), and two batches are kept in flight so
        the upload and the host-side string work of one batch overlap the kernels of the other.  Returns one list
        per page, aligned with its boxes (``None`` = empty crop, ``LineError`` = unusable region).  ``pages`` may be
        a uint8 tensor [n, H, W] in pinned host memory: its pages are then uploaded in place (no staging copy)."""
        n_pages = len(pages)
        out: List[List] = [[None] * len(b) for b in boxes_list]
        # ---- batches of whole pages
        batches, cur, cur_lines = [], [], 0
        for p in range(n_pages):
            if len(boxes_list[p]) == 0:
                continue
            cur.append(p)
            cur_lines += len(boxes_list[p])
            if cur_lines >= batch_lines:
                batches.append(cur)
                cur, cur_lines = [], 0
        if cur:
            batches.append(cur)

        as_tensor = isinstance(pages, torch.Tensor)
        if as_tensor and (pages.dtype != torch.uint8 or pages.dim() != 3):
            raise ValueError("a page tensor must be uint8 [n, H, W]")

        def launch(batch):
            arrays, ents, where, off = [], [], [], 0
            p_first = batch[0]
            layout, any_bgr = [], False
            for p in batch:
                if as_tensor:
                    shape = tuple(pages.shape[1:])
                    off = (p - p_first) * shape[0] * shape[1]   # the batch uploads pages[p_first .. p_last] as one block
                else:
                    page = np.ascontiguousarray(pages[p])
                    if page.dtype != np.uint8 or not (page.ndim == 2 or (page.ndim == 3 and page.shape[2] == 3)):
                        raise ValueError("page must be a uint8 array [H, W] (gray) or [H, W, 3] (BGR)")
                    shape = page.shape[:2]
                    arrays.append(page)
                    any_bgr |= page.ndim == 3
                    layout.append([0, off, shape[0] * shape[1], page.ndim == 3])
                ent, idx, errors = self._page_entries(shape, boxes_list[p], off)
                for i, e in errors.items():
                    out[p][i] = e
                if not as_tensor:
                    off += -(-(shape[0] * shape[1]) // 4) * 4   # gray pages start 4-byte aligned
                ents.append(ent)
                where += [(p, int(i)) for i in idx]
            ent = np.concatenate(ents) if ents else np.zeros((0, 4), np.int64)
            bgr = None
            if as_tensor:
                # zero-copy: the H2D runs straight out of the caller's (pinned) tensor
                arrays = pages[p_first:batch[-1] + 1].reshape(-1)
            else:
                # one pinned staging copy per batch; BGR pages travel as they are and become gray on the device
                arrays, raw_offs = self._stage_host(arrays, self._slot ^ 1, align=4)
                for item, ro in zip(layout, raw_offs):
                    item[0] = ro
                if any_bgr:
                    bgr = [tuple(i) for i in layout]
                else:
                    assert all(i[0] == i[1] for i in layout)    # gray only: staged offsets == gray offsets
            try:
                return (self.submit(arrays, ent, method, streaming, bgr), where, arrays, ent, bgr)
            except _lib.KiriError as e:
                return (e, where, arrays, ent, bgr)

        def finish(item):
            tk, where, arrays, ent, bgr = item
            try:
                if isinstance(tk, Exception):
                    raise tk
                res = self.collect(tk)
            except _lib.KiriError as e:
                res = self._isolate(arrays, ent, method, streaming, e, bgr)
            for (p, i), r in zip(where, res):
                out[p][i] = r

        pending = None
        for batch in batches:
            item = launch(batch)
            if pending is not None:
                finish(pending)
            pending = item
        if pending is not None:
            finish(pending)
        return out
