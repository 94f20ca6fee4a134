"""Used by __graft_entry__.smoke(): one tiny IDR + P encode on cuda:0 checked against the oracle."""
import numpy as np

import oracle
from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth


def encode_smoke():
    w, h = 320, 192
    frames = [synth.desktop(w, h, t) for t in range(3)]
    enc = oracle.RefEncoder(w, h)
    with Session(w, h, rc_mode=N.B2V_RC_CQP, crf=28) as s:
        for f in frames:
            s.submit(f)
        s.flush()
        got = s.take_frames()
    for i, f in enumerate(frames):
        ref = enc.encode_bgra(f, i == 0, qp=28)
        assert got[i].data == ref, f"encoded frame {i} differs from the oracle"
