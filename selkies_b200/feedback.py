"""Receiver-feedback bridge: RTCP feedback reaches the GPU encoder.

In the reference the RTP sender handles REMB by poking `encoder.target_bitrate` on its in-process encoder object
(src/selkies/webrtc/rtcrtpsender.py:314-322) and PLI by `_send_keyframe()` + a "pli" event that the app turns into
`request_idr_frame()` (rtcrtpsender.py:311-313, src/selkies/rtc.py:601-603).  With the pixelflux path that encoder object is not
the one producing the stream, so REMB goes nowhere.  `EncoderFeedback` is the object to hand the sender instead: it has the same
`target_bitrate` attribute (bits/s) and forwards to `ScreenCapture.update_video_bitrate(kbps)` / `request_idr_frame()`.
"""
from __future__ import annotations

import threading
import time

MIN_KBPS, MAX_KBPS = 1000, 100000          # settings.py:49 range (1..100 Mbit/s)


class EncoderFeedback:
    def __init__(self, capture, start_kbps: int = 8000, headroom: float = 0.9, min_change: float = 0.05,
                 min_interval_s: float = 0.25, min_idr_interval_s: float = 0.5, clock=time.monotonic):
        """`capture`: anything with update_video_bitrate(kbps) and request_idr_frame() — `ScreenCapture`, or a `GSTWebRTCApp`
        through `for_app()`.  `headroom`: fraction of the receiver estimate given to video (the rest is audio, RTX, FEC)."""
        self._cap, self._clock = capture, clock
        self._kbps = int(start_kbps)
        self._headroom, self._min_change = headroom, min_change
        self._min_interval, self._min_idr = min_interval_s, min_idr_interval_s
        self._t_rate = self._t_idr = -1e9
        self._lock = threading.Lock()
        self.rate_updates = self.idr_requests = 0

    # -- the attribute rtcrtpsender sets on REMB ------------------------------------------------------
    @property
    def target_bitrate(self) -> int:
        return self._kbps * 1000

    @target_bitrate.setter
    def target_bitrate(self, bps: int) -> None:
        kbps = max(MIN_KBPS, min(MAX_KBPS, int(bps * self._headroom) // 1000))
        with self._lock:
            now = self._clock()
            # decreases apply at once (congestion); increases and small moves are rate-limited so the device-side
            # controller is not re-targeted on every REMB (they arrive about once per second per receiver, but in bursts)
            down = kbps < self._kbps
            if abs(kbps - self._kbps) < self._min_change * self._kbps:
                return
            if not down and now - self._t_rate < self._min_interval:
                return
            self._kbps, self._t_rate = kbps, now
            self.rate_updates += 1
        self._cap.update_video_bitrate(kbps)

    # -- PLI / FIR -------------------------------------------------------------------------------------
    def on_pli(self, *_args) -> bool:
        """Returns True if a key frame was requested (requests closer than `min_idr_interval_s` apart collapse into one)."""
        with self._lock:
            now = self._clock()
            if now - self._t_idr < self._min_idr:
                return False
            self._t_idr = now
            self.idr_requests += 1
        self._cap.request_idr_frame()
        return True

    @classmethod
    def for_app(cls, app, **kw) -> "EncoderFeedback":
        """Bridge for a `GSTWebRTCApp`-style façade (set_video_bitrate(kbps) / send_idr())."""
        class _Shim:
            def update_video_bitrate(self, kbps): app.set_video_bitrate(int(kbps))
            def request_idr_frame(self): app.send_idr()
        return cls(_Shim(), start_kbps=int(getattr(app, "video_bitrate", 8000)), **kw)
