"""Pins oracle/csc_ref.c: BT.709 known answers, <=1 LSB vs float64, grey chroma == 128, scaling vs cv2."""
import numpy as np
import pytest

import oracle
from tests import synth

# BT.709 limited-range known answers: (R,G,B) -> (Y,Cb,Cr)
KNOWN = {
    (0, 0, 0): (16, 128, 128), (255, 255, 255): (235, 128, 128), (255, 0, 0): (63, 102, 240),
    (0, 255, 0): (173, 42, 26), (0, 0, 255): (32, 240, 118), (0, 255, 255): (188, 154, 16),
    (255, 0, 255): (78, 214, 230), (255, 255, 0): (219, 16, 138),
}


def flat(rgb, w=16, h=16):
    f = np.zeros((h, w, 4), np.uint8)
    f[..., 0], f[..., 1], f[..., 2], f[..., 3] = rgb[2], rgb[1], rgb[0], 255
    return f


@pytest.mark.parametrize("rgb,yuv", list(KNOWN.items()))
def test_known_answers(rgb, yuv):
    y, uv = oracle.csc_nv12(flat(rgb))
    assert (y == yuv[0]).all() and (uv[:, 0::2] == yuv[1]).all() and (uv[:, 1::2] == yuv[2]).all()


def float_ref(bgra):
    b, g, r = [bgra[..., i].astype(np.float64) for i in range(3)]
    kr, kb = 0.2126, 0.0722
    kg = 1 - kr - kb
    yl = kr * r + kg * g + kb * b
    y = 16 + 219 * yl / 255
    cb = 128 + 224 * (b - yl) / (2 * (1 - kb)) / 255
    cr = 128 + 224 * (r - yl) / (2 * (1 - kr)) / 255
    box = lambda a: (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) / 4
    return y, box(cb), box(cr)


def test_within_one_lsb_of_float64():
    f = synth.noise(256, 128, 3)
    y, uv = oracle.csc_nv12(f)
    fy, fcb, fcr = float_ref(f)
    assert np.abs(y - fy).max() <= 1.0
    assert np.abs(uv[:, 0::2] - fcb).max() <= 1.0
    assert np.abs(uv[:, 1::2] - fcr).max() <= 1.0
    # and it is the correctly rounded value almost everywhere
    assert (y != np.rint(fy)).mean() < 0.01


def test_grey_ramp_has_neutral_chroma():
    f = np.zeros((2, 512, 4), np.uint8)
    v = (np.arange(512) // 2).astype(np.uint8)
    f[..., 0] = f[..., 1] = f[..., 2] = v
    _, uv = oracle.csc_nv12(f)
    assert (uv == 128).all()


def test_alpha_is_ignored():
    f = synth.noise(64, 32, 5)
    g = f.copy()
    g[..., 3] = 7
    for a, b in zip(oracle.csc_nv12(f), oracle.csc_nv12(g)):
        assert (a == b).all()


def test_padding_replicates_edges():
    f = synth.noise(40, 24, 6)          # coded 48x32
    y, uv = oracle.csc_nv12(f, coded_w=48, coded_h=32)
    y0, uv0 = oracle.csc_nv12(f)
    assert (y[:24, :40] == y0).all() and (uv[:12, :40] == uv0).all()
    assert (y[:24, 40:] == y0[:, 39:40]).all() and (y[24:, :40] == y0[23:24, :]).all()


def test_scale_matches_cv2_bilinear_within_one():
    cv2 = pytest.importorskip("cv2")
    f = synth.gradient(320, 180)
    # scale the BGR planes with the oracle by converting a scaled grey image: compare luma path only
    g = f.copy()
    g[..., 1] = g[..., 0]
    g[..., 2] = g[..., 0]
    y, _ = oracle.csc_nv12(g, dst_w=192, dst_h=108)
    ref = cv2.resize(g[..., 0], (192, 108), interpolation=cv2.INTER_LINEAR).astype(np.int64)
    # grey -> Y = 16 + 219*v/255 ; invert within tolerance
    back = (y.astype(np.float64) - 16) * 255 / 219
    assert np.abs(back - ref).max() <= 2.0


def test_identity_scale_equals_unscaled():
    f = synth.noise(64, 32, 8)
    a = oracle.csc_nv12(f)
    b = oracle.csc_nv12(f, dst_w=64, dst_h=32)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
