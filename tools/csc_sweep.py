"""Sweep CSC launch shapes on the GPU box: python tools/csc_sweep.py  (writes gpurun_out/csc_sweep.json)."""
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
lib.b2v_tune_csc.argtypes = [C.c_int, C.c_int, C.c_int]
res = []
for (w, h, nres) in [(3840, 2160, 8), (7680, 4320, 4), (1920, 1080, 24)]:
    with Session(w, h, flags=N.B2V_FLAG_NO_ENCODE) as s:
        base = synth.noise(w, h, 1)
        for i in range(nres):
            s.resident_upload(i, np.roll(base, i * 7, axis=1))
        bytes_alg = w * h * 5.5
        for u in (1, 2, 3, 4):
            for block in (128, 160, 192, 256):
                for gy in (0, 148 * 2, 148 * 4):
                    lib.b2v_tune_csc(u, block, gy)
                    ms = s.bench_csc(nres, 200)
                    gbs = bytes_alg / (ms * 1e-3) / 1e9
                    res.append(dict(w=w, h=h, u=u, block=block, gy=gy, us=ms * 1e3, gbs=gbs))
                    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/csc_sweep.json", "w"), indent=1)
best = {}
for r in res:
    k = (r["w"], r["h"])
    if k not in best or r["gbs"] > best[k]["gbs"]:
        best[k] = r
print("BEST", json.dumps(list(best.values())))
