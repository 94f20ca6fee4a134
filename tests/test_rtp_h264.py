"""RTP H.264 packetiser against golden vectors produced by the UNMODIFIED reference file
(tools/make_rtp_golden.py ran src/selkies/webrtc/codecs/h264.py:165-279 in the build container)."""
import base64
import hashlib
import json
import os

import pytest

from oracle import rtp_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "rtp_h264_golden.json")))["cases"]


def check(payloads, case):
    assert [len(p) for p in payloads] == case["lens"]
    assert [hashlib.sha256(p).hexdigest()[:24] for p in payloads] == case["sha256"]
    assert [base64.b64encode(p[:8]).decode() for p in payloads] == case["first"]


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_restatement_matches_reference(i):
    check(rtp_ref.pack_access_unit(base64.b64decode(GOLD[i]["au"])), GOLD[i])


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_native_packetiser_matches_reference(i):
    from selkies_b200.rtp_h264 import H264Payloader
    check(H264Payloader().packetize(base64.b64decode(GOLD[i]["au"])), GOLD[i])


def test_golden_covers_every_packet_kind():
    kinds = set()
    for c in GOLD:
        for f in c["first"]:
            kinds.add(base64.b64decode(f)[0] & 31 if f else None)
    assert 28 in kinds and 24 in kinds and (1 in kinds or 5 in kinds)       # FU-A, STAP-A, single NAL


def test_native_vs_oracle_randomised():
    import numpy as np
    from selkies_b200.rtp_h264 import H264Payloader
    rng = np.random.default_rng(11)
    pl = H264Payloader()
    for _ in range(60):
        au = b""
        for _ in range(int(rng.integers(1, 30))):
            n = int(rng.choice([1, 2, 5, 40, 400, 1290, 1297, 1298, 1300, 1301, 2600, 9000]))
            hdr = int(rng.choice([0x01, 0x21, 0x41, 0x61, 0x65, 0x67, 0x68, 0x06]))
            body = bytes(rng.integers(4, 256, n - 1, dtype=np.uint8))
            au += (b"\x00\x00\x00\x01" if rng.integers(0, 2) else b"\x00\x00\x01") + bytes([hdr]) + body
        assert pl.packetize(au) == rtp_ref.pack_access_unit(au)
    payloads, ts = pl.pack(b"\x00\x00\x01\x65abc", 3000)
    assert payloads == [b"\x65abc"] and ts == 3000
    assert pl.pack(b"\x00\x00\x01\x41x", 1, time_base_den=30)[1] == 3000


def test_fragments_respect_mtu_and_reassemble():
    from selkies_b200.rtp_h264 import H264Payloader
    nal = bytes([0x65]) + bytes((i * 7 + 5) % 251 + 4 for i in range(10000))
    ps = H264Payloader().packetize(b"\x00\x00\x00\x01" + nal)
    assert all(len(p) <= 1300 for p in ps) and all(p[0] & 31 == 28 for p in ps)
    assert ps[0][1] & 0x80 and ps[-1][1] & 0x40 and not any(p[1] & 0xC0 for p in ps[1:-1])
    assert bytes([(ps[0][0] & 0xE0) | (ps[0][1] & 0x1F)]) + b"".join(p[2:] for p in ps) == nal
