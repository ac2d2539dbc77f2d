"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol
(no compute is issued), and the host planning logic matches the oracle / Python semantics."""
import os
import re

import numpy as np
import pytest

from kiri_ocr_b200 import _lib
from kiri_ocr_b200.config import CFG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "kiri_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(kiri_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kiri_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.kiri_version() >= 100


def test_struct_sizes_match_header_layout():
    import ctypes as C
    assert C.sizeof(_lib.KiriCropDesc) == 48
    assert C.sizeof(_lib.KiriDims) == 14 * 4
    assert C.sizeof(_lib.KiriEncLayerWeights) == 12 * 8
    assert C.sizeof(_lib.KiriDecLayerWeights) == 18 * 8


def test_missing_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.KiriError):
        _lib.require_device()


def test_target_width_matches_python_round():
    from kiri_ocr_b200.engine import target_widths
    rng = np.random.default_rng(0)
    w = rng.integers(1, 5000, 20000)
    h = rng.integers(1, 400, 20000)
    # include exact .5 cases (banker's rounding): w*48/h = k + 0.5
    w = np.concatenate([w, np.array([1, 3, 5, 7, 9, 11]) * 1]); h = np.concatenate([h, np.full(6, 96)])
    got = target_widths(w, h, 48)
    want = np.array([max(1, int(round(int(a) * (48 / float(b))))) for a, b in zip(w, h)])
    assert np.array_equal(got, want)


def test_boxes_to_entries_matches_oracle_crop():
    from kiri_ocr_b200.engine import BatchedRecognizer
    from oracle import preprocess as OP
    from tests.golden.cases import page_case
    page, boxes = page_case()
    ent, valid = BatchedRecognizer.boxes_to_entries(page.shape, boxes)
    flat = page.reshape(-1)
    for (off, pitch, w, h), ok, box in zip(ent, valid, boxes):
        roi = OP.crop_region(page, box)
        if roi is None:
            assert not ok
            continue
        assert ok
        got = np.stack([flat[off + r * pitch: off + r * pitch + w] for r in range(h)])
        raw = got if np.array_equal(got, roi) else 255 - got          # oracle applies the inversion
        assert np.array_equal(raw, roi)


def test_plan_groups_buckets_and_smem_mirror():
    from kiri_ocr_b200.engine import plan_groups, _pre_smem
    from kiri_ocr_b200 import fixtures as FX
    crops = FX.make_line_crops(200, seed=3)
    ent = np.array([(0, c.shape[1], c.shape[1], c.shape[0]) for c in crops], np.int64)
    g = plan_groups(ent, CFG(), "bucketed")
    assert set(g) <= {128, 256, 384, 512, 640} and len(g) == 5
    assert sum(len(v[0]) for v in g.values()) == 200
    for Wb, (idx, d, smem, n_strips) in g.items():
        assert (np.minimum(d["nw"], 640) <= Wb).all()
        assert smem <= 100 * 1024
        assert (d["strip_w"] <= 128).all() and n_strips == int(np.max((np.minimum(d["nw"], Wb) + d["strip_w"] - 1) // d["strip_w"]))
    gp = plan_groups(ent, CFG(), "parity")
    assert list(gp) == [640]
    # numpy mirror == C helper
    lib = _lib.load()
    for (w, h, nw, strip) in [(853, 64, 640, 640), (2000, 130, 738, 320), (50, 9, 267, 267), (300, 48, 300, 300), (1, 37, 1, 1)]:
        a = int(_pre_smem(np.array([w]), np.array([h]), np.array([nw]), 48, 640, np.array([strip]))[0])
        assert a == lib.kiri_preprocess_smem_bytes(w, h, nw, 48, 640, strip)


def test_decode_slot_table_places_every_line_once():
    """engine.decode_slot_table: every line rank appears exactly once, clusters are 16 slots, the clusters that hold the
    longest lines (the first ones) hold the fewest, and at most 15 clusters are used whenever 15 x 16 slots suffice."""
    from kiri_ocr_b200.engine import decode_slot_table
    for B in (1, 5, 16, 17, 64, 200, 225, 240, 241, 256, 1000):
        t = decode_slot_table(B)
        assert t.dtype == np.int64 and len(t) % 16 == 0
        used = t[t >= 0]
        assert np.array_equal(used, np.arange(B))                       # rank order is kept: longest lines first
        per = [(t[i:i + 16] >= 0).sum() for i in range(0, len(t), 16)]
        assert all(p >= 1 for p in per)
        assert all(a <= b for a, b in zip(per[:-2], per[1:-1]))           # non-decreasing (the last cluster takes the rest)
        for i in range(0, len(t), 16):                                   # a cluster's lines sit at the front of its slots
            n = per[i // 16]
            assert (t[i:i + n] >= 0).all() and (t[i + n:i + 16] == -1).all()
        if B <= 240:
            assert len(per) <= 15
    assert (decode_slot_table(32, n0=16) >= 0).all()                     # n0 = 16: the dense layout


def test_conv1_tensor_pipe_operand_algebra():
    """The operands csrc/conv1_tc.cu feeds the tensor pipe reproduce the normalised convolution: with u = v - 128 for a tap
    inside the image and u = -0.5 for a padded tap, sum_t (2 w_t / 255) u_t + (b + sum_t w_t / 255) equals
    sum_{inside} w_t (2 v_t / 255 - 1) + b; u and the ones column are exact in bf16, and the two-term bf16 split of the
    weights keeps 16 mantissa bits (the device accumulates the exact products in fp32)."""
    import torch
    rng = np.random.default_rng(7)
    w = (rng.standard_normal((48, 9)) / 3).astype(np.float32)
    b = (rng.standard_normal(48) * 0.1).astype(np.float32)
    v = rng.integers(0, 256, (200, 9))
    inside = rng.random((200, 9)) > 0.15
    inside[:, 4] = True                                                   # the centre tap is always inside
    x = np.where(inside, (v / 255.0 - 0.5) / 0.5, 0.0)                    # the reference's normalisation, zero padding
    want = x @ w.astype(np.float64).T + b
    u = np.where(inside, v - 128.0, -0.5)
    ub = torch.tensor(u, dtype=torch.float32).to(torch.bfloat16).double().numpy()
    assert np.array_equal(ub, u)                                          # exact in bf16
    wp = (2.0 * w.astype(np.float64) / 255.0).astype(np.float32)
    bp = (b.astype(np.float64) + w.astype(np.float64).sum(1) / 255.0).astype(np.float32)
    cols = np.concatenate([wp, bp[:, None]], 1)                           # [48, 10]: nine taps + the ones column
    hi = torch.tensor(cols).to(torch.bfloat16).float()
    lo = (torch.tensor(cols) - hi).to(torch.bfloat16).float()
    a = np.concatenate([u, np.ones((200, 1))], 1)                         # A row: [u_0 .. u_8, 1]
    got = a @ hi.double().numpy().T + a @ lo.double().numpy().T
    scale = np.abs(a) @ np.abs(cols.astype(np.float64)).T                 # sum |w' u|: the size of what is being summed
    assert float((np.abs(got - want) / scale).max()) < 2.0 ** -16          # 16 mantissa bits of weight precision
    assert float(np.abs(got - want).max()) < 2e-3                          # well under one bf16 ulp of an O(1) output


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference algorithm on the host cores, the one place outside tests/ and smoke()
    that may execute oracle/) prints ONE JSON line with the driver's keys; no GPU is touched."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "lines/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - d["config"]["lines_per_step"]) < 1e-6 * d["config"]["lines_per_step"] + 1e-6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_ctypes_signatures_agree_with_the_header():
    """Every `_lib._SIGS` entry has as many parameters as the declaration in include/kiri_b200.h, pointers where the
    header has pointers (or cudaStream_t) and scalars of the header's width elsewhere: ABI drift between the Python
    binding and the C declarations fails here, on the CPU."""
    import ctypes as C
    header = open(os.path.join(ROOT, "include", "kiri_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"//[^\n]*", "", header)
    decls = dict((m.group(1), m.group(2)) for m in re.finditer(r"\b(kiri_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    assert set(decls) == set(_lib._SIGS)
    scalar = {"int": C.c_int, "long long": C.c_longlong, "size_t": C.c_size_t, "float": C.c_float, "double": C.c_double}
    for name, params in decls.items():
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        args = _lib._SIGS[name][1]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, the binding {len(args)}"
        for p, a in zip(plist, args):
            is_ptr = "*" in p or p.startswith("cudaStream_t")
            a_ptr = a is C.c_void_p or a is C.c_char_p or (isinstance(a, type) and issubclass(a, C._Pointer))
            assert is_ptr == a_ptr, f"{name}: `{p}` bound as {a}"
            if not is_ptr:
                ty = re.sub(r"\b(const|unsigned)\b", "", p).strip()
                ty = " ".join(ty.split()[:-1])                      # drop the parameter name
                assert ty in scalar, f"{name}: unknown scalar type in `{p}`"
                assert C.sizeof(scalar[ty]) == C.sizeof(a), f"{name}: `{p}` bound as {a}"
