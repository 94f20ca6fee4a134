#!/usr/bin/env bash
# make_reference_goldens.sh — produce the reference-side golden vectors this repo cannot make itself.
#
# The legacy selkies pipeline the north_star names is  ... ! videoconvert ! x264enc ! rtph264pay ! webrtcbin  on GStreamer 1.24.12
# (pins: /root/reference/addons/gstreamer/Dockerfile:85,93; docs/component.md:318-344).  Neither GStreamer nor x264 exists in the
# build image or on the GPU box, so parity against them is UNPINNED.  Run this script on any machine that has
#   gst-launch-1.0 (1.24.x: gst-plugins-base videoconvertscale, gst-plugins-ugly x264enc)  and  python3 + numpy
# from the repo root; it writes tests/golden/reference/ (small: 4 pictures per content at 320x192, 2 at 1920x1080):
#   <name>_<w>x<h>.bgra      the synthetic inputs S1..S4 of SURVEY.md §8d (tests/synth.py), raw BGRA
#   <name>_<w>x<h>.nv12      GStreamer videoconvert output, NV12, colorimetry bt709 (limited range)
#   <name>_<w>x<h>_<kbps>.h264   x264enc tune=zerolatency speed-preset=ultrafast, constrained baseline, CBR
#   MANIFEST.json            versions (gst-launch-1.0 --version, x264enc plugin version) and the command lines used
# Commit that directory; tests/test_reference_goldens.py activates as soon as it holds a MANIFEST.json:
#   * CSC: oracle/csc_ref.c output vs the .nv12 goldens, byte for byte (reports the difference histogram per plane when not equal —
#     the rounding / chroma siting of the spec in DESIGN.md §3 is then the thing to change, in oracle/csc_ref.c AND csc.cu);
#   * H.264: PSNR of this encoder at the same bitrate vs the PSNR of the decoded x264 golden (north_star: within 0.1 dB).
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=tests/golden/reference
mkdir -p "$OUT"
command -v gst-launch-1.0 >/dev/null || { echo "gst-launch-1.0 not found: run this where GStreamer 1.24.x is installed" >&2; exit 2; }
gst-inspect-1.0 x264enc >/dev/null 2>&1 || { echo "x264enc plugin not found (gst-plugins-ugly)" >&2; exit 2; }

python3 - "$OUT" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from tests import synth
out = sys.argv[1]
gens = {"s1_bars": synth.bars, "s2_noise": lambda w, h, t: synth.noise(w, h, t), "s3_desktop": synth.desktop, "s4_gradient": synth.gradient}
for (w, h, n) in ((320, 192, 4), (1920, 1080, 2)):
    for name, g in gens.items():
        with open(os.path.join(out, f"{name}_{w}x{h}.bgra"), "wb") as f:
            for t in range(n):
                f.write(g(w, h, t).tobytes())
PY

CMDS=()
for f in "$OUT"/*.bgra; do
  base=$(basename "$f" .bgra); dims=${base##*_}; w=${dims%x*}; h=${dims#*x}
  src="filesrc location=$f ! rawvideoparse format=bgra width=$w height=$h framerate=60/1"
  c1="gst-launch-1.0 -q $src ! videoconvert ! video/x-raw,format=NV12,colorimetry=bt709 ! filesink location=$OUT/$base.nv12"
  echo "$c1"; eval "$c1"; CMDS+=("$c1")
  for kbps in 2000 8000; do
    c2="gst-launch-1.0 -q $src ! videoconvert ! video/x-raw,format=I420 ! x264enc tune=zerolatency speed-preset=ultrafast bitrate=$kbps key-int-max=600 bframes=0 byte-stream=true ! video/x-h264,profile=constrained-baseline,stream-format=byte-stream ! filesink location=$OUT/${base}_${kbps}.h264"
    echo "$c2"; eval "$c2"; CMDS+=("$c2")
  done
done

python3 - "$OUT" "$(gst-launch-1.0 --version | head -2 | tr '\n' ' ')" "$(gst-inspect-1.0 x264enc | grep -i -m1 version || true)" "${CMDS[@]}" <<'PY'
import json, sys, os
out, gst, x264 = sys.argv[1:4]
json.dump({"gst_launch": gst.strip(), "x264enc": x264.strip(), "commands": sys.argv[4:],
           "files": sorted(f for f in os.listdir(out) if not f.endswith(".json"))}, open(os.path.join(out, "MANIFEST.json"), "w"), indent=1)
PY
echo "wrote $OUT/MANIFEST.json — commit tests/golden/reference/ and run: python -m pytest tests/test_reference_goldens.py"
