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
