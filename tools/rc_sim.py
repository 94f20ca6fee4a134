"""Rate-controller behaviour on the CPU oracle (small pictures, bitrates scaled by the pixel count): achieved vs target,
worst 1-second window, QP trace.  python tools/rc_sim.py [w h n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from tests import synth


def content(name, w, h, n):
    if name == "desktop_scroll":
        return [synth.desktop(w, h, t) for t in range(n)]
    if name == "bench_cycle16":
        fr = [synth.desktop(w, h, t) for t in range(16)]
        return [fr[t % 16] for t in range(n)]
    if name == "bars_box":
        return [synth.bars(w, h, t) for t in range(n)]
    if name == "gradient_pan":
        return [synth.gradient(w, h, t) for t in range(n)]
    if name == "static_desktop":
        f = synth.desktop(w, h, 0)
        return [f] * n
    if name == "noise":
        return [synth.noise(w, h, t) for t in range(n)]
    raise ValueError(name)


def run(w, h, name, kbps, fps, n):
    frames = content(name, w, h, n)
    enc = oracle.RefEncoder(w, h)
    target = int(kbps * 1000 / fps)
    sizes, qps = [], []
    for i, f in enumerate(frames):
        au = enc.encode_bgra(f, i == 0, rc_mode=0, target_bits=target)
        sizes.append(len(au) * 8)
        qps.append(enc.last_qp)
    sizes = np.array(sizes, float)
    win = int(fps)
    tail = sizes[n // 2:]
    worst = max(sizes[i:i + win].sum() for i in range(0, n - win + 1)) / (target * win)
    worst_after = max(sizes[i:i + win].sum() for i in range(win, n - win + 1)) / (target * win) if n >= 2 * win + 1 else float("nan")
    return dict(content=name, kbps=kbps, achieved_steady=tail.mean() / target, worst_1s=worst, worst_1s_after_first_second=worst_after,
                idr_x=sizes[0] / target, qp_first=qps[0], qp_tail=(min(qps[n // 2:]), max(qps[n // 2:])), qps=qps)


if __name__ == "__main__":
    w, h, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 368, 150)
    scale = (w * h) / (3840 * 2160)
    oracle.set_threads(0)
    for name in ["desktop_scroll", "bench_cycle16", "gradient_pan", "static_desktop", "bars_box", "noise"]:
        for mbps in (8, 20, 50):
            r = run(w, h, name, mbps * 1000 * scale, 60.0, n)
            q = r.pop("qps")
            print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, "qp[::10]", q[::10], flush=True)
