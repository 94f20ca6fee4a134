"""BASELINE config 1 (1920x1080 I-only) and 4K I-only timing: python tools/intra_only.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth
for (W, H) in [(1920, 1080), (3840, 2160)]:
    frames = [synth.desktop(W, H, t) for t in range(4)]
    nb = [0]
    def onf(fp): nb[0] += fp.contents.size
    with Session(W, H, rc_mode=N.B2V_RC_CQP, crf=30, gop=1, ring_slots=4, flags=N.B2V_FLAG_TIMING, collect=False, on_frame=onf) as s:
        for i, f in enumerate(frames):
            s.resident_upload(i, f)
        for k in range(12):
            s.submit_resident(k % 4)
        s.flush(); s.reset_stats(); nb[0] = 0
        s.timer_start()
        for k in range(100):
            s.submit_resident(k % 4)
        ms = s.timer_stop()
        st = s.stats()
    print(json.dumps({"w": W, "h": H, "mode": "I-only", "fps": 100 / ms * 1e3, "bytes_per_frame": nb[0] / 100, "intra_us": st["ms_intra"] / st["n_intra"] * 1e3,
                      "cavlc_us": st["ms_cavlc"] / st["n_cavlc"] * 1e3, "slice_us": st["ms_slice"] / st["n_slice"] * 1e3, "pack_us": st["ms_pack"] / st["n_pack"] * 1e3}))
