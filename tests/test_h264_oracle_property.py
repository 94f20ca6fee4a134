"""Property test of the oracle encoder: for random small configurations every stream decodes (libavcodec) to the
encoder's own reconstruction.  Hypothesis keeps the example count small so the CPU suite stays within minutes."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from tests import synth
from tests.test_h264_oracle import run


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(w=st.sampled_from([16, 34, 64, 98, 160]), h=st.sampled_from([16, 18, 48, 82, 112]), qp=st.integers(0, 51),
       slice_rows=st.sampled_from([1, 2, 5, 100]), seed=st.integers(0, 10_000), kind=st.integers(0, 3))
def test_decode_equals_reconstruction(w, h, qp, slice_rows, seed, kind):
    rng = np.random.default_rng(seed)
    base = synth.noise(w, h, seed)
    if kind == 0:
        frames = [synth.desktop(w, h, t) for t in range(3)]
    elif kind == 1:
        frames = [synth.gradient(w, h, t) for t in range(3)]
    elif kind == 2:
        frames = [np.roll(base, (int(rng.integers(-18, 19)) * t, int(rng.integers(-18, 19)) * t), axis=(0, 1)) for t in range(3)]
    else:
        frames = [base, synth.bars(w, h, 1), synth.bars(w, h, 2)]
    run(w, h, frames, qp, slice_rows, idr_at=(0, 2) if seed % 3 == 0 else (0,))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(w=st.sampled_from([32, 66, 130, 192]), h=st.sampled_from([48, 82, 112, 146]), qp=st.sampled_from([8, 20, 28, 34, 42]),
       slice_rows=st.sampled_from([1, 2]), stripe_mult=st.integers(1, 4), seed=st.integers(0, 10_000), cbr=st.booleans(),
       paint=st.sampled_from([(0, 18), (1, 12), (2, 20)]))
def test_striped_streams_decode_independently(w, h, qp, slice_rows, stripe_mult, seed, cbr, paint):
    """Striped mode + paint-over + the temporal predictor (steady pans, still pictures, a scene change): every band's stream,
    made only of the bands the encoder flags as coded, decodes on its own to the band's rows of the reconstruction."""
    import oracle
    from oracle import avdec
    rng = np.random.default_rng(seed)
    base = synth.noise(w, h, seed)
    dx, dy = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    frames = [np.roll(base, (dy * t, dx * t), axis=(0, 1)) for t in range(4)]            # steady pan: predictor hits
    frames += [frames[-1], frames[-1], synth.desktop(w, h, 1), synth.desktop(w, h, 2)]   # still pictures, scene change, scroll
    for f in frames[1:4]:
        f[: h // 4] = frames[0][: h // 4]                                               # static top quarter
    enc = oracle.RefEncoder(w, h, slice_rows)
    stripe_rows = slice_rows * stripe_mult
    n = enc.set_stripes(stripe_rows)
    enc.set_paintover(*paint)
    streams = [[] for _ in range(n)]
    for i, f in enumerate(frames):
        au = enc.encode_bgra(f, i == 0, rc_mode=0 if cbr else 1, qp=qp, target_bits=40 * w * h // 256)
        tab = enc.stripe_table() or [(0, len(au), 1)]
        assert sum(t[1] for t in tab) == len(au)
        for k, (o, sz, coded) in enumerate(tab):
            if coded:
                streams[k].append(au[o:o + sz])
    ry, ruv = enc.recon()
    rows = stripe_rows * 16 if n > 1 else h
    for k, s in enumerate(streams):
        y0, y1 = k * rows, min(h, (k + 1) * rows)
        Y, U, V = avdec.decode_stream(s, quiet=True)[-1]
        assert np.array_equal(Y, ry[y0:y1, :w])
        assert np.array_equal(U, ruv[y0 // 2: y1 // 2, 0:w:2]) and np.array_equal(V, ruv[y0 // 2: y1 // 2, 1:w:2])
