"""CPU restatement of the reference's RTP H.264 packetiser — TEST INFRASTRUCTURE ONLY.

Follows selkies @1a9cd02b src/selkies/webrtc/codecs/h264.py (vendored aiortc 1.13.0):
  _split_bitstream   :238-263   Annex-B split on 00 00 01 (one preceding 00 dropped for 4-byte start codes)
  _packetize         :266-279   NAL > PACKET_MAX -> FU-A, else try STAP-A aggregation
  _packetize_fu_a    :165-201   near-equal fragments of <= 1298 payload bytes, S/E bits
  _packetize_stap_a  :204-235   up to 9 NALs, 1297 bytes budget, F/NRI = max over aggregated NALs
PINNED: tests/golden/rtp_h264_golden.json was produced by executing the unmodified reference file
(tools/make_rtp_golden.py); tests/test_rtp_h264.py checks this restatement and the native packetiser against it.
"""
from __future__ import annotations

import math
from struct import pack

PACKET_MAX = 1300          # h264.py:58
NAL_TYPE_FU_A, NAL_TYPE_STAP_A = 28, 24


def split_bitstream(buf: bytes):
    i = 0
    while True:
        i = buf.find(b"\x00\x00\x01", i)
        if i == -1:
            return
        i += 3
        start = i
        i = buf.find(b"\x00\x00\x01", i)
        if i == -1:
            yield buf[start:]
            return
        yield buf[start:i - 1] if buf[i - 1] == 0 else buf[start:i]


def packetize_fu_a(data: bytes):
    avail = PACKET_MAX - 2
    size = len(data) - 1
    n = math.ceil(size / avail)
    larger, per = size % n, size // n
    ind = (data[0] & 0xE0) | NAL_TYPE_FU_A
    nal = data[0] & 0x1F
    out, off, hdr = [], 1, bytes([ind, nal | 0x80])
    while off < len(data):
        take = per + 1 if larger > 0 else per
        larger -= 1 if larger > 0 else 0
        payload = data[off:off + take]
        off += take
        if off == len(data):
            hdr = bytes([ind, nal | 0x40])
        out.append(hdr + payload)
        hdr = bytes([ind, nal])
    return out


def packetize(nals):
    out = []
    it = iter(nals)
    cur = next(it, None)
    while cur is not None:
        if len(cur) > PACKET_MAX:
            out += packetize_fu_a(cur)
            cur = next(it, None)
            continue
        avail, count, hdr, payload, nalu = PACKET_MAX - 3, 0, NAL_TYPE_STAP_A | (cur[0] & 0xE0), b"", cur
        while nalu is not None and len(nalu) <= avail and count < 9:
            hdr |= nalu[0] & 0x80
            nri = nalu[0] & 0x60
            if hdr & 0x60 < nri:
                hdr = hdr & 0x9F | nri
            avail -= 2 + len(nalu)
            count += 1
            payload += pack("!H", len(nalu)) + nalu
            nalu = next(it, None)
        if count == 0:
            nalu = next(it, None)
        out.append(cur if count <= 1 else bytes([hdr]) + payload)
        cur = nalu
    return out


def pack_access_unit(au: bytes):
    return packetize(split_bitstream(au))
