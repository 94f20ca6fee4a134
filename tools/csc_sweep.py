"""Compare the CSC kernels on the GPU box: python tools/csc_sweep.py  (writes gpurun_out/csc_sweep.json).
b2v_tune_csc(u, block, gy): gy = -1 LDG fast path (u units per thread, `block` threads), -2 / -3 TMA path with 1 / 2 CTAs per SM."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth

lib = N.lib()
res = []
for (w, h, nres) in [(3840, 2160, 8), (7680, 4320, 4), (1920, 1080, 24)]:
    with Session(w, h, flags=N.B2V_FLAG_NO_ENCODE) as s:
        base = synth.noise(w, h, 1)
        for i in range(nres):
            s.resident_upload(i, np.roll(base, i * 7, axis=1))
        bytes_alg = w * h * 5.5
        for name, args in [("ldg_u2_b160", (2, 160, -1)), ("ldg_u4_b256", (4, 256, -1)), ("tma_1cta", (2, 160, -2)), ("tma_2cta", (2, 160, -3))]:
            lib.b2v_tune_csc(*args)
            ms = s.bench_csc(nres, 200)
            msb = s.bench_csc_burst(nres, 200)
            res.append(dict(w=w, h=h, kernel=name, us_event_pair=ms * 1e3, us_burst=msb * 1e3, gbs_event=bytes_alg / (ms * 1e-3) / 1e9, gbs_burst=bytes_alg / (msb * 1e-3) / 1e9))
            print(res[-1], flush=True)
lib.b2v_tune_csc(2, 160, -3)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/csc_sweep.json", "w"), indent=1)
