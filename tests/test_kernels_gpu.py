"""Kernel-level parity through the C ABI on a real B200 (-m gpu).

Checkers: the CPU oracle (oracle/), torch fp32 ops for floating-point kernels, and the library's
own CUDA-core reference GEMM.  Tolerances are stated per test; integer outputs are bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from kiri_ocr_b200 import _lib  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    _lib.require_device()
    return _lib.load()


def dev(t):
    return t.cuda().contiguous()


def sync():
    torch.cuda.synchronize()


# --------------------------------------------------------------------------- CTC greedy
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_ctc_greedy_matches_oracle(lib, dtype):
    from oracle import decode as OD
    rng = np.random.default_rng(0)
    B, T, C_, ld = 9, 160, 204, 208
    x = rng.normal(0, 2.0, (B, T, ld)).astype(np.float32)
    x[0, :, 0] += 50.0                         # all blank -> empty line
    x[1, 10:40, 7] += 50.0                     # long repeat
    x[2, ::2, 1] += 50.0                       # pad id (1) interleaved
    x[3, :, 2] += 50.0                         # one symbol for the whole line
    x[4, 5:9, 203] += 50.0                     # last class
    xt = torch.from_numpy(x)
    if dtype == "bf16":
        xt = xt.to(torch.bfloat16)
    xd = dev(xt)
    ids = torch.full((B, T), -1, dtype=torch.int32, device="cuda")
    n_ids = torch.zeros(B, dtype=torch.int32, device="cuda")
    conf = torch.zeros(B, dtype=torch.float32, device="cuda")
    fids = torch.zeros((B, T), dtype=torch.int32, device="cuda")
    fprob = torch.zeros((B, T), dtype=torch.float32, device="cuda")
    _lib.check(lib.kiri_ctc_greedy(xd.data_ptr(), _lib.DTYPE_BF16 if dtype == "bf16" else _lib.DTYPE_F32, B, T, C_, ld,
                                   ids.data_ptr(), n_ids.data_ptr(), conf.data_ptr(), fids.data_ptr(),
                                   fprob.data_ptr(), _lib.stream_ptr()))
    sync()
    xr = xt.float().numpy()[:, :, :C_]
    for b in range(B):
        best, collapsed, cf, length = OD.ctc_greedy(xr[b])
        assert np.array_equal(fids[b].cpu().numpy(), best.astype(np.int32)), b
        n = int(n_ids[b])
        assert n == length == len(collapsed)
        assert np.array_equal(ids[b, :n].cpu().numpy(), collapsed)
        assert abs(float(conf[b]) - cf) < 2e-6 * max(1.0, cf) + 1e-6          # fp32 mean of soft-max maxima
    assert int(n_ids[0]) == 0 and int(n_ids[3]) == 1


# --------------------------------------------------------------------------- preprocess
def run_preprocess(lib, crops_or_pages, boxes, Wb, img_h=48, want_norm=False, smem_cap=200 * 1024, strip_cap=128):
    """crops_or_pages: list of 2-D uint8 arrays; boxes: list of (page_idx, x, y, w, h) already clamped."""
    offs, total = [], 0
    for p in crops_or_pages:
        offs.append(total)
        total += p.size
    buf = np.zeros(total + 16, np.uint8)
    for p, o in zip(crops_or_pages, offs):
        buf[o:o + p.size] = p.reshape(-1)
    descs = (_lib.KiriCropDesc * len(boxes))()
    smem, max_strips = 0, 1
    for i, (pi, x, y, w, h) in enumerate(boxes):
        page = crops_or_pages[pi]
        nw = max(1, int(round(w * (img_h / float(h)))))
        wout = min(nw, Wb)
        strip = min(wout, strip_cap)
        while lib.kiri_preprocess_smem_bytes(w, h, nw, img_h, Wb, strip) > smem_cap and strip > 32:
            strip = max(32, (strip // 2 + 31) // 32 * 32)
        need = lib.kiri_preprocess_smem_bytes(w, h, nw, img_h, Wb, strip)
        smem = max(smem, need)
        max_strips = max(max_strips, (wout + strip - 1) // strip)
        d = descs[i]
        d.src_offset = offs[pi] + y * page.shape[1] + x
        d.pitch, d.w, d.h, d.nw, d.Wb, d.strip_w = page.shape[1], w, h, nw, Wb, strip
        d.out_offset = i * img_h * Wb
    src = torch.from_numpy(buf).cuda()
    dd = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).cuda()
    planes = torch.zeros((len(boxes), img_h, Wb), dtype=torch.uint8, device="cuda")
    norm = torch.zeros((len(boxes), img_h, Wb), dtype=torch.bfloat16, device="cuda") if want_norm else None
    sums = torch.zeros(len(boxes), dtype=torch.int64, device="cuda")
    _lib.check(lib.kiri_preprocess_pack(src.data_ptr(), dd.data_ptr(), len(boxes), img_h, smem, max_strips,
                                        planes.data_ptr(), _lib.ptr(norm), sums.data_ptr(), _lib.stream_ptr()))
    sync()
    return planes.cpu().numpy(), (norm.float().cpu().numpy() if want_norm else None)


def test_preprocess_golden_crops_bit_exact(lib, golden):
    from oracle import preprocess as OP
    from tests.golden.cases import golden_crops
    crops = golden_crops()
    boxes = [(i, 0, 0, c.shape[1], c.shape[0]) for i, c in enumerate(crops)]
    planes, norm = run_preprocess(lib, crops, boxes, 640, want_norm=True)
    for i, c in enumerate(crops):
        want = OP.resize_keep_ratio_pad(OP.crop_region(c, (0, 0, c.shape[1], c.shape[0]), 0))
        assert np.array_equal(planes[i], want), i
        nref = torch.from_numpy(OP.normalise(want)).to(torch.bfloat16).float().numpy()
        assert np.array_equal(norm[i], nref), i
    assert np.array_equal(planes[0], golden["hard/0/plane"])


def test_preprocess_random_shapes_and_buckets(lib):
    from oracle import preprocess as OP
    rng = np.random.default_rng(5)
    for Wb in (128, 256, 384, 512, 640):
        crops = []
        for _ in range(12):
            h = int(rng.integers(1, 140))
            w = int(rng.integers(1, 2200))
            a = rng.integers(0, 256, (h, w), dtype=np.uint8)
            if rng.random() < 0.3:
                a = (a // 3).astype(np.uint8)                  # dark -> invert branch
            crops.append(a)
        boxes = [(i, 0, 0, c.shape[1], c.shape[0]) for i, c in enumerate(crops)]
        planes, _ = run_preprocess(lib, crops, boxes, Wb)
        for i, c in enumerate(crops):
            want = OP.resize_keep_ratio_pad(OP.crop_region(c, (0, 0, c.shape[1], c.shape[0]), 0), 48, Wb)
            assert np.array_equal(planes[i], want), (Wb, i, c.shape)


def test_preprocess_page_boxes_and_strips(lib):
    """Crops taken straight from a page (unaligned offsets, pitch != w) and a tall crop that
    forces the multi-strip path."""
    from oracle import preprocess as OP
    from tests.golden.cases import page_case
    page, boxes = page_case()
    tall = np.random.default_rng(1).integers(0, 256, (700, 900), dtype=np.uint8)
    clamped = []
    want = []
    for (x, y, w, h) in boxes:
        x1, y1 = max(0, x - 5), max(0, y - 5)
        x2, y2 = min(page.shape[1], x + w + 5), min(page.shape[0], y + h + 5)
        if x2 <= x1 or y2 <= y1:
            continue
        clamped.append((0, x1, y1, x2 - x1, y2 - y1))
        want.append(OP.preprocess_region(page, (x, y, w, h)))
    clamped.append((1, 3, 7, 801, 650))
    want.append(OP.resize_keep_ratio_pad(OP.crop_region(tall[7:657, 3:804], (0, 0, 801, 650), 0)))
    planes, _ = run_preprocess(lib, [page, tall], clamped, 640, smem_cap=96 * 1024)
    for i, wv in enumerate(want):
        assert np.array_equal(planes[i], wv), i


# --------------------------------------------------------------------------- GEMM (tcgen05)
def gemm(lib, a, w, bias, epi, resid=None, ln=None):
    M, K = a.shape
    N = w.shape[0]
    f32_out = epi in (_lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32, _lib.EPI_BIAS_RESID_LN)
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if f32_out else torch.bfloat16, device="cuda")
    out2 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda") if epi == _lib.EPI_BIAS_RESID_LN else None
    if resid is not None:
        out.copy_(resid)
    g, b = (ln if ln is not None else (None, None))
    _lib.check(lib.kiri_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, N, K, epi, out.data_ptr(),
                                  out.data_ptr() if resid is not None else 0, _lib.ptr(g), _lib.ptr(b),
                                  _lib.ptr(out2), _lib.stream_ptr()), "kiri_gemm_bf16")
    sync()
    return out, out2


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (256, 256, 64), (384, 768, 256), (300, 1024, 256),
                                   (1000, 256, 1024), (160, 208, 256), (77, 416, 256), (40960, 256, 256)])
def test_gemm_bias_f32_vs_reference_kernel(lib, M, N, K):
    torch.manual_seed(M + N + K)
    a = dev(torch.randn(M, K).to(torch.bfloat16))
    w = dev((torch.randn(N, K) / K ** 0.5).to(torch.bfloat16))
    bias = dev(torch.randn(N))
    ref = torch.empty((M, N), dtype=torch.float32, device="cuda")
    _lib.check(lib.kiri_gemm_ref(a.data_ptr(), w.data_ptr(), M, N, K, ref.data_ptr(), _lib.stream_ptr()))
    out, _ = gemm(lib, a, w, bias, _lib.EPI_BIAS_F32)
    want = ref + bias
    tref = a.float() @ w.float().t() + bias
    assert float((want - tref).abs().max()) < 1e-3
    err = float((out - want).abs().max())
    assert err < 2e-3, f"max abs err {err}"          # fp32 accumulation order only


def test_gemm_epilogues(lib):
    torch.manual_seed(3)
    M, N, K = 520, 256, 256
    a = dev(torch.randn(M, K).to(torch.bfloat16))
    w = dev((torch.randn(N, K) / 16).to(torch.bfloat16))
    bias = dev(torch.randn(N) * 0.5)
    acc = a.float() @ w.float().t() + bias
    out, _ = gemm(lib, a, w, bias, _lib.EPI_BIAS_BF16)
    assert float((out.float() - acc).abs().max()) < 0.03           # bf16 output rounding (|v| < 8)
    out, _ = gemm(lib, a, w, bias, _lib.EPI_BIAS_SILU_BF16)
    assert float((out.float() - F.silu(acc)).abs().max()) < 0.03
    out, _ = gemm(lib, a, w, bias, _lib.EPI_BIAS_GELU_BF16)
    assert float((out.float() - F.gelu(acc)).abs().max()) < 0.03
    resid = dev(torch.randn(M, N))
    out, _ = gemm(lib, a, w, bias, _lib.EPI_BIAS_RESID_F32, resid=resid)
    assert float((out - (acc + resid)).abs().max()) < 2e-3
    g, b = dev(torch.rand(N) + 0.5), dev(torch.randn(N) * 0.1)
    out, out2 = gemm(lib, a, w, bias, _lib.EPI_BIAS_RESID_LN, resid=resid, ln=(g, b))
    x = acc + resid
    assert float((out - x).abs().max()) < 2e-3
    want = F.layer_norm(x, (N,), g, b, 1e-5)
    assert float((out2.float() - want).abs().max()) < 0.04


# --------------------------------------------------------------------------- conv (implicit GEMM)
@pytest.mark.parametrize("cin,cout,sh,sw,IH,IW,n", [
    (64, 96, 2, 2, 48, 640, 2), (96, 160, 2, 2, 24, 320, 2), (160, 256, 2, 1, 12, 160, 3),
    (64, 96, 2, 2, 48, 256, 1), (64, 96, 2, 2, 48, 384, 2), (96, 160, 2, 2, 24, 192, 1),
    (160, 256, 2, 1, 12, 128, 2), (160, 256, 2, 1, 12, 32, 5), (96, 160, 2, 2, 24, 64, 3)])
def test_conv3x3_vs_torch(lib, cin, cout, sh, sw, IH, IW, n):
    torch.manual_seed(cin + IW)
    x = dev((torch.randn(n, IH, IW, cin)).to(torch.bfloat16))                 # NHWC
    w = dev((torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)).to(torch.bfloat16))
    bias = dev(torch.randn(cout) * 0.2)
    wk = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    OH, OW = (IH - 1) // sh + 1, (IW - 1) // sw + 1
    out = torch.full((n, OH, OW, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.kiri_conv3x3_bf16(x.data_ptr(), wk.data_ptr(), bias.data_ptr(), n, IH, IW, cin, cout, sh, sw,
                                     out.data_ptr(), 0, _lib.stream_ptr()), "kiri_conv3x3_bf16")
    sync()
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, (sh, sw), 1)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert not torch.isnan(out.float()).any()
    assert float(err.max()) < 0.03, f"max err {float(err.max())} at {torch.nonzero(err == err.max())[0].tolist()}"


@pytest.mark.parametrize("IW,n", [(640, 2), (128, 3), (384, 1)])
def test_conv2_on_dense_48_channels_equals_padded_64(lib, IW, n):
    """conv2 reads conv1's DENSE 48-channel activation: the tensor map's inner extent is 48, the box 64, and the TMA
    unit zero-fills channels 48..63 in shared memory.  Must be bit-identical to the same conv on a 64-channel tensor
    whose last 16 channels are stored zeros (the round-1 layout), and match torch."""
    torch.manual_seed(IW + n)
    IH, cout = 48, 96
    x48 = dev(torch.randn(n, IH, IW, 48).to(torch.bfloat16))
    x64 = torch.zeros(n, IH, IW, 64, dtype=torch.bfloat16, device="cuda")
    x64[..., :48] = x48
    w = (torch.randn(cout, 48, 3, 3) / (3 * 48 ** 0.5)).to(torch.bfloat16)
    w64 = torch.zeros(cout, 3, 3, 64, dtype=torch.bfloat16)
    w64[..., :48] = w.permute(0, 2, 3, 1)
    # garbage in the padded weight columns must not matter either: the activations there are zero-filled
    wk = dev(w64.reshape(cout, 9 * 64))
    bias = dev(torch.randn(cout) * 0.2)
    outs = []
    for x, cm in ((x48, 48), (x64, 0)):
        out = torch.full((n, IH // 2, IW // 2, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.kiri_conv3x3_bf16(x.data_ptr(), wk.data_ptr(), bias.data_ptr(), n, IH, IW, 64, cout, 2, 2, out.data_ptr(),
                                         cm, _lib.stream_ptr()), "kiri_conv3x3_bf16")
        sync()
        outs.append(out)
    assert not torch.isnan(outs[0].float()).any()
    assert torch.equal(outs[0], outs[1])
    ref = F.silu(F.conv2d(x48.float().permute(0, 3, 1, 2), dev(w).float(), bias, (2, 2), 1)).permute(0, 2, 3, 1)
    assert float((outs[0].float() - ref).abs().max()) < 0.03


@pytest.mark.parametrize("n,W", [(3, 256), (2, 640), (1, 128)])
def test_conv1_vs_torch(lib, n, W):
    """conv1 (packed fp32 FMAs, dense 48-channel NHWC output) against torch in float64."""
    torch.manual_seed(W)
    H = 48
    planes = torch.randint(0, 256, (n, H, W), dtype=torch.uint8)
    planes[0, :4] = 0
    planes[0, 4:8] = 255
    w = torch.randn(48, 9) / 3
    b = torch.randn(48) * 0.1
    out = torch.full((n, H, W, 48), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.kiri_conv1(dev(planes).data_ptr(), w.data_ptr(), b.data_ptr(), n, H, W, out.data_ptr(), _lib.stream_ptr()))
    sync()
    x = (planes.float() / 255.0 - 0.5) / 0.5
    ref = F.silu(F.conv2d(x[:, None].double(), w.view(48, 1, 3, 3).double(), b.double(), 1, 1)).permute(0, 2, 3, 1)
    o = out.float().cpu()
    assert not torch.isnan(o).any()
    err = (o.double() - ref).abs()
    # the only error left is the bf16 rounding of the output (2^-9 relative) and tanh.approx (2^-11)
    assert float((err / (ref.abs() + 0.05)).max()) < 6e-3, float((err / (ref.abs() + 0.05)).max())
    assert float(err.max()) < 0.03


@pytest.mark.parametrize("groups", [[(3, 256)], [(2, 640), (1, 128), (2, 384)]])
def test_conv1_tensor_pipe_vs_torch_and_cuda_core_form(lib, groups):
    """conv1 on the tensor pipe (csrc/conv1_tc.cu: exact bf16 operands u = v - 128 / -0.5 for padded taps, weights split
    into two bf16 terms, bias and normalisation folded into the operand) against torch in float64 - same bound as the
    CUDA-core form - and against the CUDA-core form itself (at most one bf16 ulp apart, on a small fraction of outputs)."""
    torch.manual_seed(11)
    H = 48
    w = torch.randn(48, 9) / 3
    b = torch.randn(48) * 0.1
    planes = [torch.randint(0, 256, (n, H, W), dtype=torch.uint8) for n, W in groups]
    planes[0][0, :4] = 0
    planes[0][0, 4:8] = 255
    dplanes = [dev(p) for p in planes]
    outs = [torch.full((n, H, W, 48), float("nan"), dtype=torch.bfloat16, device="cuda") for n, W in groups]
    outs_cc = [torch.full((n, H, W, 48), float("nan"), dtype=torch.bfloat16, device="cuda") for n, W in groups]
    k = len(groups)
    args = lambda o: ((C.c_void_p * k)(*[p.data_ptr() for p in dplanes]), (C.c_void_p * k)(*[m.data_ptr() for m in o]),
                      (C.c_int * k)(*[n for n, _ in groups]), (C.c_int * k)(*[W for _, W in groups]), k, w.data_ptr(), b.data_ptr(), H,
                      _lib.stream_ptr())
    _lib.check(lib.kiri_conv1_tc_multi(*args(outs)))
    _lib.check(lib.kiri_conv1_multi(*args(outs_cc)))
    sync()
    for pl, o, occ in zip(planes, outs, outs_cc):
        x = (pl.float() / 255.0 - 0.5) / 0.5
        ref = F.silu(F.conv2d(x[:, None].double(), w.view(48, 1, 3, 3).double(), b.double(), 1, 1)).permute(0, 2, 3, 1)
        of = o.float().cpu()
        assert not torch.isnan(of).any()
        err = (of.double() - ref).abs()
        assert float((err / (ref.abs() + 0.05)).max()) < 6e-3, float((err / (ref.abs() + 0.05)).max())
        assert float(err.max()) < 0.03
        d = (of - occ.float().cpu()).abs()
        assert float((d / (ref.abs().float() + 0.05)).max()) < 9e-3          # one bf16 ulp (2^-8 relative) at most
        assert float((d > 0).float().mean()) < 0.05


def test_conv1_multi_groups_equal_single_launches(lib):
    """One launch for several width groups == one launch per group, bit for bit."""
    torch.manual_seed(5)
    H = 48
    w = torch.randn(48, 9) / 3
    b = torch.randn(48) * 0.1
    shapes = [(2, 128), (3, 384), (1, 640)]
    planes = [dev(torch.randint(0, 256, (n, H, W), dtype=torch.uint8)) for n, W in shapes]
    single = []
    for (n, W), pl in zip(shapes, planes):
        o = torch.empty((n, H, W, 48), dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.kiri_conv1(pl.data_ptr(), w.data_ptr(), b.data_ptr(), n, H, W, o.data_ptr(), _lib.stream_ptr()))
        single.append(o)
    multi = [torch.full((n, H, W, 48), float("nan"), dtype=torch.bfloat16, device="cuda") for n, W in shapes]
    k = len(shapes)
    _lib.check(lib.kiri_conv1_multi((C.c_void_p * k)(*[p.data_ptr() for p in planes]), (C.c_void_p * k)(*[m.data_ptr() for m in multi]),
                                    (C.c_int * k)(*[n for n, _ in shapes]), (C.c_int * k)(*[W for _, W in shapes]), k,
                                    w.data_ptr(), b.data_ptr(), H, _lib.stream_ptr()))
    sync()
    for a, m in zip(single, multi):
        assert torch.equal(a, m)


# --------------------------------------------------------------------------- norms / attention
def test_pool_pos_ln_and_layernorm(lib):
    torch.manual_seed(1)
    n, RH, T, D = 3, 6, 96, 256
    act = dev(torch.randn(n, RH, T, D).to(torch.bfloat16))
    pos = dev(torch.randn(T, D))
    g0, b0, g1, b1 = (dev(torch.rand(D) + 0.5), dev(torch.randn(D) * 0.1), dev(torch.rand(D) + 0.5), dev(torch.randn(D) * 0.1))
    x = torch.empty(n * T, D, device="cuda")
    a = torch.empty(n * T, D, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.kiri_pool_pos_ln(act.data_ptr(), pos.data_ptr(), n, RH, T, D, g0.data_ptr(), b0.data_ptr(),
                                    g1.data_ptr(), b1.data_ptr(), x.data_ptr(), a.data_ptr(), _lib.stream_ptr()))
    sync()
    pooled = act.float().mean(1) + pos
    xr = F.layer_norm(pooled, (D,), g0, b0, 1e-5).reshape(n * T, D)
    assert float((x - xr).abs().max()) < 1e-4
    ar = F.layer_norm(xr, (D,), g1, b1, 1e-5)
    assert float((a.float() - ar).abs().max()) < 0.04
    y = torch.empty_like(x)
    yb = torch.empty_like(a)
    z = torch.empty_like(a)
    _lib.check(lib.kiri_layernorm(x.data_ptr(), n * T, D, g1.data_ptr(), b1.data_ptr(), y.data_ptr(), yb.data_ptr(),
                                  g0.data_ptr(), b0.data_ptr(), z.data_ptr(), _lib.stream_ptr()))
    sync()
    yr = F.layer_norm(x, (D,), g1, b1, 1e-5)
    assert float((y - yr).abs().max()) < 1e-4
    assert float((z.float() - F.layer_norm(yr, (D,), g0, b0, 1e-5)).abs().max()) < 0.04


@pytest.mark.parametrize("T", [32, 64, 96, 128, 160])
def test_encoder_attention_vs_torch(lib, T):
    torch.manual_seed(T)
    n, heads, D = 3, 8, 256
    qkv = dev((torch.randn(n * T, 3 * D)).to(torch.bfloat16))
    out = torch.full((n * T, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.kiri_encoder_attention(qkv.data_ptr(), out.data_ptr(), n, T, heads, D, 0, _lib.stream_ptr()))
    sync()
    q, k, v = (t.reshape(n, T, heads, 32).transpose(1, 2) for t in qkv.float().split(D, dim=1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n * T, D)
    assert float((out.float() - ref).abs().max()) < 0.03
    # masked variant: keys beyond kv_len are ignored
    kv_len = torch.tensor([T, T // 2, 5], dtype=torch.int32, device="cuda")
    _lib.check(lib.kiri_encoder_attention(qkv.data_ptr(), out.data_ptr(), n, T, heads, D, kv_len.data_ptr(), _lib.stream_ptr()))
    sync()
    for i, L in enumerate(kv_len.tolist()):
        r = F.scaled_dot_product_attention(q[i:i + 1], k[i:i + 1, :, :L], v[i:i + 1, :, :L]).transpose(1, 2).reshape(T, D)
        assert float((out[i * T:(i + 1) * T].float() - r).abs().max()) < 0.03


def test_encoder_attention_multi_group_one_launch(lib):
    """All width groups of a concatenated token stream in ONE launch (group order != length order)."""
    import ctypes as C
    torch.manual_seed(7)
    heads, D = 8, 256
    lines, Ts = [2, 3, 0, 1, 4, 2], [64, 160, 96, 32, 128, 96]
    M = sum(n * T for n, T in zip(lines, Ts))
    qkv = dev(torch.randn(M, 3 * D).to(torch.bfloat16))
    kv = []
    for n, T in zip(lines, Ts):
        kv += [max(1, T - 7 * i) for i in range(n)]
    kv_len = torch.tensor(kv, dtype=torch.int32, device="cuda")
    gl, gt = (C.c_int * len(lines))(*lines), (C.c_int * len(Ts))(*Ts)
    for masked in (False, True):
        out = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.kiri_encoder_attention_multi(qkv.data_ptr(), out.data_ptr(), gl, gt, len(lines), heads, D,
                                                    kv_len.data_ptr() if masked else 0, _lib.stream_ptr()))
        sync()
        r0, li = 0, 0
        for n, T in zip(lines, Ts):
            for i in range(n):
                blk = qkv[r0:r0 + T].float()
                q, k, v = (t.reshape(1, T, heads, 32).transpose(1, 2) for t in blk.split(D, dim=1))
                L = kv[li] if masked else T
                ref = F.scaled_dot_product_attention(q, k[:, :, :L], v[:, :, :L]).transpose(1, 2).reshape(T, D)
                assert float((out[r0:r0 + T].float() - ref).abs().max()) < 0.03, (T, i, masked)
                r0 += T
                li += 1
        assert not torch.isnan(out.float()).any()


@pytest.mark.parametrize("M,FF,with_ln", [(128, 1024, True), (26080, 1024, True), (544, 1024, False), (40960, 1024, True),
                                           (96, 256, True)])
def test_encoder_block_fused_vs_three_gemms_and_torch(lib, M, FF, with_ln):
    """encoder_block.cu (out_proj + LN + FFN + LN in one kernel) against (a) the three tcgen05 GEMM launches it
    replaces and (b) a plain torch fp32 reference of the same ops on the same bf16-rounded operands."""
    torch.manual_seed(M + FF)
    D = 256
    o = dev((torch.randn(M, D) * 0.7).to(torch.bfloat16))
    x0 = dev(torch.randn(M, D))
    wo = dev((torch.randn(D, D) / 16).to(torch.bfloat16))
    w1 = dev((torch.randn(FF, D) / 16).to(torch.bfloat16))
    w2 = dev((torch.randn(D, FF) / 32).to(torch.bfloat16))
    bo, b1, b2 = dev(torch.randn(D) * 0.1), dev(torch.randn(FF) * 0.1), dev(torch.randn(D) * 0.1)
    g1, h1 = dev(1 + 0.1 * torch.randn(D)), dev(0.1 * torch.randn(D))
    g2, h2 = dev(1 + 0.1 * torch.randn(D)), dev(0.1 * torch.randn(D))
    # (b) torch fp32 reference
    xm = x0 + o.float() @ wo.float().T + bo
    a2 = F.layer_norm(xm, (D,), g1, h1, 1e-5).to(torch.bfloat16).float()
    h = F.gelu(a2 @ w1.float().T + b1).to(torch.bfloat16).float()
    xr = xm + h @ w2.float().T + b2
    ar = F.layer_norm(xr, (D,), g2, h2, 1e-5)
    # (a) the three launches
    x3 = x0.clone()
    a3 = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
    hb = torch.zeros(M, FF, dtype=torch.bfloat16, device="cuda")
    a_fin = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
    s = _lib.stream_ptr()
    _lib.check(lib.kiri_gemm_bf16(o.data_ptr(), wo.data_ptr(), bo.data_ptr(), M, D, D, 5, x3.data_ptr(), x3.data_ptr(),
                                  g1.data_ptr(), h1.data_ptr(), a3.data_ptr(), s))
    _lib.check(lib.kiri_gemm_bf16(a3.data_ptr(), w1.data_ptr(), b1.data_ptr(), M, FF, D, 2, hb.data_ptr(), 0, 0, 0, 0, s))
    _lib.check(lib.kiri_gemm_bf16(hb.data_ptr(), w2.data_ptr(), b2.data_ptr(), M, D, FF, 5, x3.data_ptr(), x3.data_ptr(),
                                  g2.data_ptr(), h2.data_ptr(), a_fin.data_ptr(), s))
    # fused
    x = x0.clone()
    a = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    for rep in range(2):                      # twice: the second run exercises warm barriers / PDL back to back
        x.copy_(x0)
        _lib.check(lib.kiri_encoder_block(o.data_ptr(), x.data_ptr(), a.data_ptr() if with_ln else 0, wo.data_ptr(), bo.data_ptr(),
                                          w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), g1.data_ptr(), h1.data_ptr(),
                                          g2.data_ptr() if with_ln else 0, h2.data_ptr() if with_ln else 0, M, FF, s))
    sync()
    assert not torch.isnan(x).any()
    ex_ref = float((x - xr).abs().max())
    ex_3 = float((x - x3).abs().max())
    assert ex_ref < 0.03, ex_ref                  # bf16 operands / hidden activation, fp32 accumulate
    assert ex_3 < 0.02, ex_3                      # same arithmetic, different fp32 summation order
    if with_ln:
        ea_ref = float((a.float() - ar).abs().max())
        ea_3 = float((a.float() - a_fin.float()).abs().max())
        assert ea_ref < 0.06, ea_ref
        assert ea_3 < 0.05, ea_3


def test_pack_records_matches_numpy(lib):
    """Fixed-stride exchange records {n_ids, conf bits, ids[T]} from the token-major CTC output."""
    rng = np.random.default_rng(3)
    T = 160
    lens = np.array([160, 32, 96, 64, 128], np.int32)
    row0 = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    ids = rng.integers(2, 200, int(lens.sum())).astype(np.int32)
    n_ids = np.array([0, 32, 17, 5, 100], np.int32)
    conf = rng.random(5).astype(np.float32)
    rec = torch.full((5, 2 + T), -7, dtype=torch.int32, device="cuda")
    d = lambda a: torch.from_numpy(a).cuda()
    t_ids, t_n, t_c, t_r = d(ids), d(n_ids), d(conf), d(row0)
    _lib.check(lib.kiri_pack_records(t_ids.data_ptr(), t_n.data_ptr(), t_c.data_ptr(), t_r.data_ptr(), 5, T, rec.data_ptr(),
                                     _lib.stream_ptr()))
    sync()
    got = rec.cpu().numpy()
    for b in range(5):
        assert got[b, 0] == n_ids[b] and got[b, 1:2].view(np.float32)[0] == conf[b]
        want = np.zeros(T, np.int32)
        want[:n_ids[b]] = ids[row0[b]:row0[b] + n_ids[b]]
        assert np.array_equal(got[b, 2:], want)


def test_encoder_block_soak_back_to_back():
    """1000 back-to-back launches (PDL on) of the fused encoder tail at the token counts of the bench (26 080 bucketed,
    40 960 parity), one full wave (18 944) and a single tile, each twice from the same input: finite and bit-identical.
    A second process repeats it with the clock64 phase accounting on (KIRI_GEMM_TIMING=1), the configuration of the
    crash log that round 1 committed as profiles/r01_encoder_block_phase_cycles_v2.txt."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import eb_soak
    lib = _lib.load()
    for M in (128, 18944, 26080, 40960):
        eb_soak.soak(lib, M, 1000)
        eb_soak.soak(lib, M, 1000, affine=False)             # the engine's variant (LayerNorm affines folded into the weights)
    eb_soak.soak(lib, 26080, 300, with_ln=False)
    env = dict(os.environ, KIRI_GEMM_TIMING="1")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "eb_soak.py"), "300"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("soak ok") == 8
    # third process: the CHECKED build (make -C csrc checked) - device-side invariants of the ring / barrier phases / TMEM base
    # (ring units issued == consumed, tiles per role, ...) trap with a message instead of corrupting memory
    assert os.path.exists(_lib.CHECKED_LIB_PATH), "libkiri_b200_checked.so is missing: run __graft_entry__.build()"
    env = dict(os.environ, KIRI_B200_LIB=_lib.CHECKED_LIB_PATH)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "eb_soak.py"), "300"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("soak ok") == 8 and "libkiri_b200_checked.so" in r.stdout and "KIRI_CHECKED" not in r.stdout


def test_bgr_to_gray_bit_exact_vs_cv2(lib):
    """GPU page ingest: the BGR -> gray kernel equals cv2.cvtColor(COLOR_BGR2GRAY) byte for byte (core.py:762-766)."""
    import cv2
    rng = np.random.default_rng(7)
    for shape in ((37, 53), (480, 641), (2339, 1654), (1, 1), (3, 2)):
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        img[0, : min(shape[1], 256), 0] = np.arange(min(shape[1], 256))
        want = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        src = dev(torch.from_numpy(img.reshape(-1)))
        out = torch.zeros(shape[0] * shape[1] + 8, dtype=torch.uint8, device="cuda")
        _lib.check(lib.kiri_bgr_to_gray(src.data_ptr(), shape[0] * shape[1], out.data_ptr(), _lib.stream_ptr()))
        sync()
        got = out.cpu().numpy()
        assert np.array_equal(got[: want.size].reshape(shape), want), shape
        assert not got[want.size:].any()                      # nothing written past the image
