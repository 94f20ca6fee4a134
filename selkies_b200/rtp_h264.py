"""RTP H.264 payloader backed by the native packetiser (include/b2video.h: b2v_rtp_h264_packetize).

Mirror of the part of the reference's `H264Encoder` that handles pre-encoded video — `pack(packet)` ->
`(list[bytes], timestamp)` (src/selkies/webrtc/codecs/h264.py:331-335, used by RTCRtpSender._next_encoded_frame,
rtcrtpsender.py:362) — i.e. the "rtph264pay" of the legacy pipeline.  Payloads are byte-identical to the
reference's `_packetize(_split_bitstream(...))` (h264.py:238-279).
"""
from __future__ import annotations

import ctypes as C

from . import _native as N

PACKET_MAX = 1300            # h264.py:58
VIDEO_CLOCK_RATE = 90000     # webrtc/codecs/__init__.py:109


class H264Payloader:
    def __init__(self, packet_max: int = PACKET_MAX):
        self._lib = N.lib()
        self.packet_max = packet_max
        self._cap = 0
        self._out = None
        self._lens = None

    def _reserve(self, au_size: int):
        need = au_size + au_size // 600 + 64
        if need > self._cap:
            self._cap = max(need, 2 * self._cap, 1 << 16)
            self._out = C.create_string_buffer(self._cap)
            self._lens = (C.c_int32 * (self._cap // 64 + 16))()

    def packetize(self, au: bytes) -> list[bytes]:
        """Annex-B access unit -> list of RTP payloads."""
        au = bytes(au)
        self._reserve(len(au))
        n = C.c_int32(0)
        rc = self._lib.b2v_rtp_h264_packetize(au, len(au), self.packet_max, self._out, self._cap, self._lens, len(self._lens), C.byref(n))
        if rc != 0:
            raise N.B2VError(rc, "b2v_rtp_h264_packetize failed (empty NAL or buffer too small)")
        mv = memoryview(self._out)
        out, off = [], 0
        for i in range(n.value):
            ln = self._lens[i]
            out.append(bytes(mv[off:off + ln]))
            off += ln
        return out

    def pack(self, data: bytes, pts: int, time_base_den: int = VIDEO_CLOCK_RATE) -> tuple[list[bytes], int]:
        """`H264Encoder.pack` for raw bytes: pts in units of 1/time_base_den seconds -> 90 kHz timestamp."""
        ts = pts if time_base_den == VIDEO_CLOCK_RATE else (pts * VIDEO_CLOCK_RATE) // time_base_den
        return self.packetize(data), ts
