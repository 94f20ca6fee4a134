"""Hand-off bridge between the encoder callback and the RTP sender (host logic; the payloader itself is covered by test_rtp_h264)."""
import asyncio

from selkies_b200.rtc_bridge import PipelineBridge, VideoBridge, VideoSample


def test_bridge_keeps_only_the_newest_sample():
    async def run():
        b = PipelineBridge()
        for i in range(5):
            await b.set_data(i)
        assert await b.get_data() == 4 and b.dropped == 4
        getter = asyncio.ensure_future(b.get_data())
        await asyncio.sleep(0)
        assert not getter.done()
        await b.set_data("x")
        assert await getter == "x"
    asyncio.run(run())


def test_consume_data_makes_a_packet_like_sample():
    class FakePayloader:
        def pack(self, data, pts, den):
            return [data[:3], data[3:]], pts * 90000 // den

    async def run():
        vb = VideoBridge(FakePayloader())
        raw = bytearray(b"\x00\x00\x00\x01\x65abcdef")
        await vb.consume_data(memoryview(raw), 3000, "video", is_keyframe=True)
        raw[:] = b"\xff" * len(raw)                        # the callback's buffer is reused by the encoder: the sample owns a copy
        await vb.consume_data(b"", 6000, "video")          # empty buffers are ignored (rtc.py:410)
        await vb.consume_data(b"zz", 6000, "audio")        # not ours
        s = await vb.video_pipeline_bridge.get_data()
        assert isinstance(s, VideoSample) and bytes(s) == b"\x00\x00\x00\x01\x65abcdef" and s.pts == s.dts == 3000
        assert s.time_base.denominator == 90000 and s.is_keyframe and len(s) == 11
        payloads, ts = vb.pack(s)
        assert payloads == [b"\x00\x00\x00", b"\x01\x65abcdef"] and ts == 3000
    asyncio.run(run())


def test_bridge_with_the_native_payloader_matches_the_reference_packetiser():
    """An oracle-encoded access unit through consume_data -> pack: same payloads as the restated reference packetiser."""
    import numpy as np
    import oracle
    from oracle import rtp_ref
    from tests import synth

    enc = oracle.RefEncoder(320, 192, 1)
    au = enc.encode_bgra(synth.noise(320, 192, 4), True, rc_mode=1, qp=24, target_bits=0)

    async def run():
        vb = VideoBridge()
        await vb.consume_data(au, 90000 // 60 * 7, "video", True)
        return vb.pack(await vb.video_pipeline_bridge.get_data())
    payloads, ts = asyncio.run(run())
    assert ts == 90000 // 60 * 7
    assert payloads == rtp_ref.pack_access_unit(au) and max(len(p) for p in payloads) <= 1300
