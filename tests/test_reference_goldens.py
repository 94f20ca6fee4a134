"""Closes the reference pin when someone supplies the goldens: tests/golden/reference/ is filled by
tools/make_reference_goldens.sh on a machine with GStreamer 1.24.x (videoconvert, x264enc — the legacy selkies CPU pipeline,
/root/reference/addons/gstreamer/Dockerfile:85,93).  Without MANIFEST.json every test here is skipped and the parity status
stays "unpinned" (DESIGN.md §2)."""
import glob
import json
import os
import re

import numpy as np
import pytest

import oracle

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference")
MANIFEST = os.path.join(HERE, "MANIFEST.json")
needs_goldens = pytest.mark.skipif(not os.path.exists(MANIFEST), reason="no reference goldens: run tools/make_reference_goldens.sh where GStreamer 1.24.x exists")


def cases(ext):
    out = []
    for p in sorted(glob.glob(os.path.join(HERE, "*" + ext))):
        m = re.search(r"_(\d+)x(\d+)", os.path.basename(p))
        if m:
            out.append((p, int(m.group(1)), int(m.group(2))))
    return out


def test_manifest_or_readme_present():
    assert os.path.exists(MANIFEST) or os.path.exists(os.path.join(HERE, "README.md"))


@needs_goldens
def test_csc_matches_videoconvert_bit_exact():
    """north_star: NV12 bytes bit-exact against GStreamer videoconvert on the same BGRA input."""
    assert json.load(open(MANIFEST))["files"]
    report = []
    for path, w, h in cases(".nv12"):
        bgra = np.fromfile(path[:-5] + ".bgra", np.uint8).reshape(-1, h, w, 4)
        nv12 = np.fromfile(path, np.uint8).reshape(-1, h * 3 // 2, w)
        for i, f in enumerate(bgra):
            y, uv = oracle.csc_nv12(f)
            gy, guv = nv12[i][:h], nv12[i][h:]
            dy, duv = y.astype(int) - gy.astype(int), uv.astype(int) - guv.astype(int)
            if dy.any() or duv.any():
                report.append((os.path.basename(path), i, {int(v): int(c) for v, c in zip(*np.unique(dy, return_counts=True))},
                               {int(v): int(c) for v, c in zip(*np.unique(duv, return_counts=True))}))
    assert not report, "CSC spec (DESIGN.md §3) differs from videoconvert; difference histograms (Y, CbCr) per picture: " + repr(report[:4])


@needs_goldens
@pytest.mark.gpu
def test_psnr_within_a_tenth_of_a_db_of_x264():
    """north_star: the encoded bitstream decodes to within 0.1 dB PSNR of x264enc at the same bitrate/preset."""
    from oracle import avdec
    from selkies_b200 import _native as N
    from selkies_b200.session import Session
    worst = []
    for path, w, h in cases(".h264"):
        kbps = int(re.search(r"_(\d+)\.h264$", path).group(1))
        src = re.sub(r"_\d+\.h264$", ".bgra", path)
        frames = list(np.fromfile(src, np.uint8).reshape(-1, h, w, 4))
        ref_dec = avdec.decode_stream([open(path, "rb").read()], quiet=True)
        with Session(w, h, fps=60.0, rc_mode=N.B2V_RC_CBR, bitrate_kbps=kbps) as s:
            for f in frames:
                s.submit(f)
            s.flush()
            got = s.take_frames()
        dec = avdec.decode_stream([g.data for g in got], quiet=True)
        for i, f in enumerate(frames[: min(len(ref_dec), len(dec))]):
            sy, _ = oracle.csc_nv12(f)
            worst.append((avdec.psnr(ref_dec[i][0], sy) - avdec.psnr(dec[i][0], sy), os.path.basename(path), i))
    assert worst and max(worst)[0] <= 0.1, sorted(worst, reverse=True)[:4]
