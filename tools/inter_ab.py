"""A/B the k_inter_mb variants on the GPU box: B2V_INTER_VARIANT=0|1 python tools/inter_ab.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth
W, H = 3840, 2160
frames = [synth.desktop(W, H, t) for t in range(8)]
SR = int(os.environ.get('SLICE_ROWS', '1'))
nbytes = [0]
def onf(fp): nbytes[0] += fp.contents.size
with Session(W, H, rc_mode=N.B2V_RC_CQP, crf=30, slice_rows=SR, ring_slots=4, flags=N.B2V_FLAG_TIMING, collect=False, on_frame=onf) as s:
    for i, f in enumerate(frames):
        s.resident_upload(i, f)
    for k in range(48):
        s.submit_resident(k % 8)
    s.flush(); s.reset_stats()
    s.timer_start()
    for k in range(320):
        s.submit_resident(k % 8)
    ms = s.timer_stop()
    st = s.stats()
print(json.dumps({"slice_rows": SR, "bytes_per_frame": nbytes[0] / 368, "fps": 320 / ms * 1e3,
                  "inter_us": st["ms_inter"] / max(1, st["n_inter"]) * 1e3, "cavlc_us": st["ms_cavlc"] / st["n_cavlc"] * 1e3,
                  "slice_us": st["ms_slice"] / st["n_slice"] * 1e3, "pack_us": st["ms_pack"] / st["n_pack"] * 1e3}))
