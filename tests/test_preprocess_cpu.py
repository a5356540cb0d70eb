"""Oracle of the demo's pre-processing (oracle/preprocess.py) against (1) the committed golden vectors produced with the
real dependency (Pillow, tests/golden/make_golden_preprocess.py) and (2) the live Pillow when it is importable."""
import os

import numpy as np
import pytest

from oracle import preprocess as pp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


def test_oracle_reproduces_pillow_golden_vectors_bit_exactly():
    g = np.load(GOLDEN)
    for key in g.files:
        if not key.startswith("scale_"):
            continue
        got = pp.net_input_from_image(g["image"], float(key[6:]))
        assert got.shape == g[key].shape, key
        assert np.array_equal(got, g[key]), key


@pytest.mark.parametrize("h,w,scale", [(37, 53, 0.5), (40, 64, 1.5), (33, 47, 0.73), (50, 50, 2.0), (61, 35, 0.31), (48, 64, 1.0),
                                       (1, 1, 1.0), (2, 3, 0.2), (180, 320, 0.85)])
def test_resize_matches_live_pillow(h, w, scale):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ow, oh = max(1, int(w * scale)), max(1, int(h * scale))
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(pp.pillow_bilinear_resize_u8(img, ow, oh), want)


def test_scale_one_is_replicate_pad_minus_mean():
    img = np.random.default_rng(1).integers(0, 256, (13, 21, 3), dtype=np.uint8)
    x = pp.net_input_from_image(img, 1.0)
    assert x.shape == (3, 16, 24)
    assert np.array_equal(x[:, :13, :21], img.transpose(2, 0, 1).astype(np.float32) - pp.MEAN.reshape(3, 1, 1).astype(np.float32))
    assert np.array_equal(x[:, 13:, :21], np.broadcast_to(x[:, 12:13, :21], (3, 3, 21)))      # rows below replicate the last row
    assert np.array_equal(x[:, :, 21:], np.broadcast_to(x[:, :, 20:21], (3, 16, 3)))


def test_coefficient_rows_sum_to_one_in_fixed_point():
    for n_in, n_out in ((109, 54), (100, 150), (784, 392), (64, 64)):
        _, bounds, kk = pp.resample_coeffs(n_in, n_out)
        assert np.all(np.abs(kk.sum(axis=1) - (1 << pp.PRECISION_BITS)) <= kk.shape[1])
        assert np.all(bounds[:, 0] >= 0) and np.all(bounds[:, 0] + bounds[:, 1] <= n_in)
