"""The oracle's integer resample against the installed Pillow (the reference's own dependency,
model.py:323).  CPU only; skipped when Pillow is absent."""
import numpy as np
import pytest

from oracle import preprocess as OP

Image = pytest.importorskip("PIL.Image")


def _ref(a, img_h=48, img_w=640):
    img = Image.fromarray(a)
    iw, ih = img.size
    nw = max(1, int(round(iw * (img_h / float(ih)))))
    img = img.resize((nw, img_h), Image.BILINEAR)
    if nw >= img_w:
        return np.array(img.crop((0, 0, img_w, img_h)))
    canvas = Image.new("L", (img_w, img_h), 128)
    canvas.paste(img, (0, 0))
    return np.array(canvas)


def test_random_shapes_bit_exact():
    rng = np.random.default_rng(0)
    for _ in range(120):
        h = int(rng.integers(1, 140))
        w = int(rng.integers(1, 2600))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(OP.resize_keep_ratio_pad(a), _ref(a)), (h, w)


@pytest.mark.parametrize("shape", [(48, 640), (48, 1), (1, 1), (96, 1280), (24, 320), (47, 641), (49, 639),
                                   (200, 30), (5, 4000)])
def test_edge_shapes_bit_exact(shape):
    rng = np.random.default_rng(sum(shape))
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(OP.resize_keep_ratio_pad(a), _ref(a))


def test_other_model_sizes():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (70, 500), dtype=np.uint8)
    for W in (128, 256, 384, 512):
        assert np.array_equal(OP.resize_keep_ratio_pad(a, 48, W), _ref(a, 48, W))


def test_normalise_matches_reference_op_order():
    import torch
    p = np.arange(256, dtype=np.uint8).reshape(16, 16)
    t = torch.from_numpy(p).float() / 255.0
    t = (t - 0.5) / 0.5
    assert np.array_equal(OP.normalise(p), t.numpy())
