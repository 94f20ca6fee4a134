"""GPU parity, striped mode ("x264enc-striped", CaptureSettings.h264_fullframe = False, selkies.py:3219): the CUDA encoder's
stripes vs oracle/h264_ref.c, bit-exact, through the C-ABI; every stripe stream decodes on its own (libavcodec)."""
import numpy as np
import pytest

import oracle
from oracle import avdec
from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth

pytestmark = pytest.mark.gpu


def run_both(w, h, frames, stripe_rows, slice_rows, *, qp=28, rc_mode=N.B2V_RC_CQP, kbps=0, fps=30.0, idr_at=(0,), header_mode=N.B2V_HDR_NONE, paint=(0, 18)):
    enc = oracle.RefEncoder(w, h, slice_rows)
    enc.set_stripes(stripe_rows)
    enc.set_paintover(*paint)
    target = int(kbps * 1000 / fps) if kbps else 0
    ref = []                       # per picture: [(y_start, bytes)] of the coded bands
    for i, f in enumerate(frames):
        au = enc.encode_bgra(f, i in idr_at, rc_mode=1 if rc_mode == N.B2V_RC_CQP else 0, qp=qp, target_bits=target)
        ref.append([(k * stripe_rows * 16, au[o:o + sz]) for k, (o, sz, c) in enumerate(enc.stripe_table()) if c])
    with Session(w, h, rc_mode=rc_mode, crf=qp, bitrate_kbps=kbps or 8000, fps=fps, slice_rows=slice_rows, gop=-1,
                 stripe_rows=stripe_rows, header_mode=header_mode, paintover_trigger_frames=paint[0], paintover_crf=paint[1]) as s:
        for i, f in enumerate(frames):
            if i in idr_at and i > 0:
                s.flush()
                s.request_idr()
            s.submit(f)
        s.flush()
        got = s.take_frames()
        grec = s.recon()
    return got, ref, grec, enc.recon()


def group(got, n_pictures):
    per = [[] for _ in range(n_pictures)]
    for g in got:
        per[g.frame_id].append(g)
    return per


def frames_with_static_top(w, h, n):
    frames = [synth.desktop(w, h, t) for t in range(n)]
    for f in frames[1:]:
        f[:64] = frames[0][:64]
    return frames


@pytest.mark.parametrize("w,h,stripe_rows,slice_rows", [(320, 200, 4, 1), (320, 200, 4, 2), (320, 200, 6, 3), (192, 112, 1, 1), (640, 360, 8, 1), (320, 200, 4, 0), (640, 360, 6, 0)])
def test_stripes_bit_exact(w, h, stripe_rows, slice_rows):
    frames = frames_with_static_top(w, h, 5)
    got, ref, grec, rrec = run_both(w, h, frames, stripe_rows, slice_rows, idr_at=(0, 3))
    per = group(got, len(frames))
    for i, (gp, rp) in enumerate(zip(per, ref)):
        assert [(g.y_start, g.data) for g in gp] == rp, f"picture {i}"
        for g in gp:
            assert g.height == min(h, g.y_start + stripe_rows * 16) - g.y_start and g.is_key == (i in (0, 3))
    assert np.array_equal(grec[0], rrec[0]) and np.array_equal(grec[1], rrec[1])
    if (w, stripe_rows) == (320, 4):
        assert all(g.y_start >= 64 for g in per[1])           # the static band is not sent
    # every stripe is a stream of its own
    for y0 in sorted({g.y_start for g in got}):
        st = [g.data for g in got if g.y_start == y0]
        Y, U, V = avdec.decode_stream(st, quiet=True)[-1]
        y1 = min(h, y0 + stripe_rows * 16)
        assert np.array_equal(Y, grec[0][y0:y1, :w]) and np.array_equal(U, grec[1][y0 // 2: y1 // 2, 0:w:2])


def test_stripes_subpel_motion_stays_inside_the_band():
    """A vertical pan: vectors that would leave the band see the band's edge padding instead, exactly as each band's decoder does."""
    cv2 = pytest.importorskip("cv2")
    w, h, rows = 320, 192, 3
    rng = np.random.default_rng(7)
    big = cv2.GaussianBlur(rng.normal(0, 1, (h + 96, w + 96, 3)).astype(np.float32), (0, 0), 4) * 700 + 128
    frames = []
    for t in range(4):
        a = cv2.warpAffine(big, np.float32([[1, 0, -(24 + 0.75 * t)], [0, 1, -(24 + 3.25 * t)]]), (w, h), flags=cv2.INTER_CUBIC)
        frames.append(np.dstack([np.clip(a, 0, 255).astype(np.uint8), np.full((h, w), 255, np.uint8)]))
    got, ref, grec, rrec = run_both(w, h, frames, rows, 1, qp=26)
    per = group(got, len(frames))
    for gp, rp in zip(per, ref):
        assert [(g.y_start, g.data) for g in gp] == rp
    assert np.array_equal(grec[0], rrec[0]) and np.array_equal(grec[1], rrec[1])
    for y0 in range(0, h, rows * 16):
        Y, _, _ = avdec.decode_stream([g.data for g in got if g.y_start == y0], quiet=True)[-1]
        assert np.array_equal(Y, grec[0][y0: y0 + rows * 16, :w])


def test_stripes_cbr_and_pixelflux_header():
    w, h, rows = 320, 200, 4
    frames = frames_with_static_top(w, h, 8)
    got, ref, grec, rrec = run_both(w, h, frames, rows, 1, rc_mode=N.B2V_RC_CBR, kbps=900, header_mode=N.B2V_HDR_PIXELFLUX)
    per = group(got, len(frames))
    for i, (gp, rp) in enumerate(zip(per, ref)):
        assert [(g.y_start, g.data[10:]) for g in gp] == rp, f"picture {i}"
        for g in gp:
            d = g.data
            assert d[0] == 0x04 and d[1] == (1 if i == 0 else 0) and int.from_bytes(d[2:4], "big") == i
            assert int.from_bytes(d[4:6], "big") == g.y_start and int.from_bytes(d[6:8], "big") == w and int.from_bytes(d[8:10], "big") == g.height
    assert np.array_equal(grec[0], rrec[0])
    assert len({g.qp for g in got}) > 1


def test_stripe_rows_must_be_a_multiple_of_slice_rows():
    with pytest.raises(Exception):
        Session(320, 200, rc_mode=N.B2V_RC_CQP, crf=28, slice_rows=2, stripe_rows=3).close()


def test_screen_capture_striped_mode():
    """CaptureSettings.h264_fullframe = False -> stripes with the 10-byte header, one decoder per y_start (selkies-ws-core.js:3183-3216)."""
    import threading
    from selkies_b200.pixelflux_compat import ArraySource, CaptureSettings, ScreenCapture
    w, h = 320, 200
    frames = frames_with_static_top(w, h, 6)
    cs = CaptureSettings()
    cs.capture_width, cs.capture_height, cs.target_fps = w, h, 120.0
    cs.h264_fullframe, cs.h264_crf, cs.h264_stripe_rows = False, 28, 4
    seen, done = [], threading.Event()

    def cb(result_ptr, _):
        r = result_ptr.contents
        seen.append(bytes(r.data[:r.size]))
        if int.from_bytes(seen[-1][2:4], "big") >= 5 and int.from_bytes(seen[-1][4:6], "big") == 192:
            done.set()

    cap = ScreenCapture(ArraySource(frames, loop=False))
    cap.start_capture(cs, cb)
    assert done.wait(20)
    cap.stop_capture()
    by_y = {}
    for d in seen:
        assert d[0] == 0x04
        by_y.setdefault(int.from_bytes(d[4:6], "big"), []).append(d[10:])
    assert sorted(by_y) == [0, 64, 128, 192]
    assert len(by_y[0]) == 1 and len(by_y[64]) == 6          # static top stripe: key frame only
    for y0, st in by_y.items():
        Y, _, _ = avdec.decode_stream(st, quiet=True)[-1]
        assert Y.shape == (min(h, y0 + 64) - y0, w)


def test_ws_video_channel_carries_stripes_from_the_native_thread():
    """ScreenCapture (striped) -> WsVideoChannel.on_stripe on the native output thread -> sender task -> per-stripe decoders."""
    import asyncio
    from selkies_b200.pixelflux_compat import ArraySource, CaptureSettings, ScreenCapture, StripeCallback
    from selkies_b200.ws_video import WsVideoChannel
    w, h = 320, 200
    frames = frames_with_static_top(w, h, 6)

    async def run():
        loop = asyncio.get_running_loop()
        sent = []

        async def send(b):
            sent.append(b)
        ch = WsVideoChannel(send, loop, queue_depth=256, fps=120.0)
        sender = asyncio.ensure_future(ch.run_sender())
        cs = CaptureSettings()
        cs.capture_width, cs.capture_height, cs.target_fps = w, h, 120.0
        cs.h264_fullframe, cs.h264_crf, cs.h264_stripe_rows = False, 28, 4
        cap = ScreenCapture(ArraySource(frames, loop=False))
        await loop.run_in_executor(None, cap.start_capture, cs, StripeCallback(ch.on_stripe))
        for _ in range(400):
            await asyncio.sleep(0.01)
            if ch.last_sent_id >= 5 and ch.queue.empty():
                break
        await loop.run_in_executor(None, cap.stop_capture)
        await asyncio.sleep(0.05)
        await ch.queue.join()
        sender.cancel()
        ch.on_ack(f"CLIENT_FRAME_ACK {ch.last_sent_id}")
        assert ch.evaluate_gate() and ch.dropped == 0
        return sent
    sent = asyncio.run(run())
    by_y = {}
    for d in sent:
        assert d[0] == 0x04
        by_y.setdefault(int.from_bytes(d[4:6], "big"), []).append(d[10:])
    assert sorted(by_y) == [0, 64, 128, 192] and len(by_y[0]) == 1
    for y0, st in by_y.items():
        Y, _, _ = avdec.decode_stream(st, quiet=True)[-1]
        assert Y.shape == (min(h, y0 + 64) - y0, w)
