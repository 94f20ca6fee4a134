"""Striped mode of the oracle (pixelflux h264_fullframe = False, "x264enc-striped", selkies.py:3219): every band is an
independent H.264 stream.  Pinned two ways on CPU: libavcodec decodes each band's stream on its own to exactly the band's rows of
the encoder reconstruction, and a band's bytes equal what a separate full-frame encoder instance produces for the cropped band."""
import numpy as np
import pytest

import oracle
from oracle import avdec
from tests import synth

W, H = 320, 200          # 13 macroblock rows, coded height 208 (bottom crop 8)


def band_frames(n=4):
    frames = [synth.desktop(W, H, t) for t in range(n)]
    for f in frames[1:]:
        f[:64] = frames[0][:64]          # the top 64 rows never change: their band(s) must be dropped after the IDR
    return frames


def encode_striped(frames, stripe_rows, slice_rows, qp=28, idr_at=(0,)):
    enc = oracle.RefEncoder(W, H, slice_rows)
    n = enc.set_stripes(stripe_rows)
    streams = [[] for _ in range(n)]
    for i, f in enumerate(frames):
        au = enc.encode_bgra(f, i in idr_at, rc_mode=1, qp=qp, target_bits=0)
        tab = enc.stripe_table()
        assert sum(t[1] for t in tab) == len(au) and [t[0] for t in tab] == list(np.cumsum([0] + [t[1] for t in tab[:-1]]))
        for k, (o, sz, coded) in enumerate(tab):
            if coded:
                streams[k].append(au[o:o + sz])
    return enc, streams


@pytest.mark.parametrize("stripe_rows,slice_rows", [(4, 1), (4, 2), (6, 3), (1, 1), (5, 5)])
def test_each_band_decodes_alone_to_its_rows(stripe_rows, slice_rows):
    enc, streams = encode_striped(band_frames(), stripe_rows, slice_rows)
    ry, ruv = enc.recon()
    for k, st in enumerate(streams):
        r0 = k * stripe_rows * 16
        r1 = min(H, r0 + stripe_rows * 16)
        assert len(st) == (1 if r1 <= 64 else 4)          # a band inside the static rows is sent once; the rest every picture
        Y, U, V = avdec.decode_stream(st, quiet=True)[-1]
        assert Y.shape == (r1 - r0, W)
        assert np.array_equal(Y, ry[r0:r1, :W])
        assert np.array_equal(U, ruv[r0 // 2: r1 // 2, 0:W:2]) and np.array_equal(V, ruv[r0 // 2: r1 // 2, 1:W:2])


def test_band_equals_an_independent_encoder_on_the_cropped_band():
    frames = band_frames(5)
    stripe_rows = 4
    enc, streams = encode_striped(frames, stripe_rows, 1, idr_at=(0, 3))
    for k, st in enumerate(streams):
        r0 = k * stripe_rows * 16
        r1 = min(H, r0 + stripe_rows * 16)
        solo = oracle.RefEncoder(W, r1 - r0, 1)
        got = []
        for i, f in enumerate(frames):
            au = solo.encode_bgra(np.ascontiguousarray(f[r0:r1]), i in (0, 3), rc_mode=1, qp=28, target_bits=0)
            got.append(au)
        if k == 0:      # static band: the solo encoder codes all-skip P pictures the striped one drops
            assert st == [got[0], got[3]]
        else:
            assert st == got


def test_stripe_rows_must_align_with_slices():
    enc = oracle.RefEncoder(W, H, 2)
    with pytest.raises(ValueError):
        enc.set_stripes(3)
    assert enc.set_stripes(0) == 1 and enc.set_stripes(13) == 1     # off / one band = full frame
