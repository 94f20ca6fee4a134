"""A/B of two builds of libb2video.so on the same box: the bench's side legs (S2 noise, S4 gradient pan, headline) device-timed.
B2V_LIB=<path> python tools/ab_legs.py   (run once per library; each process loads one build)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                     # noqa: E402
from selkies_b200 import _native as N            # noqa: E402
from tests import synth                          # noqa: E402

W, H = 3840, 2160
cbr = dict(fps=60.0, rc_mode=N.B2V_RC_CBR, bitrate_kbps=20000, ring_slots=16)
out = {"lib": os.environ.get("B2V_LIB", "in-tree")}
for rep in range(2):
    out[f"s2_{rep}"] = round(bench.resident_leg([synth.noise(W, H, 100 + t) for t in range(4)], 96, 0, warm=16, **cbr)["value"])
    out[f"s4_{rep}"] = round(bench.resident_leg([synth.gradient(W, H, t) for t in range(16)], 128, 0, warm=32, **cbr)["value"])
    out[f"head_{rep}"] = round(bench.resident_leg([synth.desktop(W, H, t) for t in range(16)], 512, 0, warm=48, **cbr)["value"])
print(json.dumps(out))
