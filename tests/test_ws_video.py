"""WS-mode video channel: queue, sender, ACK-driven gate (host logic, no GPU, no sockets)."""
import asyncio

from selkies_b200.pixelflux_compat import _Result, _ResultPtr
from selkies_b200.ws_video import WsVideoChannel, frames_behind, gate_is_open


def test_gate_arithmetic():
    assert frames_behind(10, 4) == 6 and frames_behind(3, 65533) == 6                  # u16 wrap
    assert gate_is_open(500, -1, 60, 0, 99)                                            # nothing acknowledged yet
    assert gate_is_open(0, 7, 60, 0, 0)
    assert gate_is_open(200, 100, 60, 0, 0.1)                                          # 100 frames behind <= 120 allowed
    assert not gate_is_open(230, 100, 60, 0, 0.1)                                      # 130 > 120
    assert gate_is_open(230, 100, 60, 200, 0.1)                                        # 200 ms RTT credits 12 frames
    assert not gate_is_open(230, 100, 60, 40, 0.1)                                     # RTT below 50 ms is not credited
    assert not gate_is_open(110, 100, 60, 0, 4.5)                                      # client silent for > 4 s
    assert gate_is_open(40000, 100, 60, 0, 0.1)                                        # implausible gap: ignored
    assert gate_is_open(130, 100, 0, 0, 0.1) and not gate_is_open(230, 100, 0, 0, 0.1)  # fps 0 falls back to 60


def stripe(payload: bytes, frame_id: int):
    r = _Result()
    r.data, r.size, r.frame_id = memoryview(payload), len(payload), frame_id
    return _ResultPtr(r)


def test_channel_end_to_end():
    async def run():
        loop = asyncio.get_running_loop()
        t = [100.0]
        out = []

        async def send(b):
            out.append(b)
        ch = WsVideoChannel(send, loop, queue_depth=4, fps=60.0, clock=lambda: t[0])
        sender = asyncio.ensure_future(ch.run_sender())
        for i in range(1, 4):
            ch.on_stripe(stripe(b"\x04\x00" + bytes([0, i]) + b"stripe", i))
        ch.on_stripe(None)
        await asyncio.sleep(0.01)
        await ch.queue.join()
        assert [b[3] for b in out] == [1, 2, 3] and ch.last_sent_id == 3 and ch.sent == 3
        t[0] += 0.030
        assert ch.on_ack("CLIENT_FRAME_ACK 2") and not ch.on_ack("CLIENT_FRAME_ACK x") and not ch.on_ack("HELLO 1")
        assert abs(ch.rtt_ms - 30.0) < 1e-6 and ch.acked_id == 2
        assert ch.evaluate_gate() is True
        # the client falls 130 frames behind: gate closes, stripes are consumed but not sent
        ch.last_sent_id = 132
        assert ch.evaluate_gate() is False
        ch.on_stripe(stripe(b"\x04\x00\x00\x85late", 133))
        await asyncio.sleep(0.01)
        await ch.queue.join()
        assert len(out) == 3 and ch.last_sent_id == 132
        ch.on_ack("CLIENT_FRAME_ACK 130")
        assert ch.evaluate_gate() is True
        # queue overflow drops, never blocks the native thread
        sender.cancel()
        for i in range(8):
            ch._offer(b"x", 200 + i)
        assert ch.dropped == 4
    asyncio.run(run())


def test_jpeg_prefix():
    async def run():
        got = []

        async def send(b):
            got.append(b)
        ch = WsVideoChannel(send, asyncio.get_running_loop(), jpeg=True)
        task = asyncio.ensure_future(ch.run_sender())
        ch.on_stripe(stripe(b"JFIF", 1))
        await asyncio.sleep(0.01)
        await ch.queue.join()
        task.cancel()
        assert got == [b"\x03\x00JFIF"]
    asyncio.run(run())


def test_gate_keeps_state_before_first_frame_and_resets_stall_timer():
    """selkies.py:1239 (`if server_id == 0: continue`) leaves the gate as it is; :1225 / :1236 restart the stall timer."""
    from selkies_b200.ws_video import gate_decision, gate_is_open
    assert gate_decision(0, 5, 60.0, 0.0, 0.0) == (None, False)
    assert gate_is_open(0, 5, 60.0, 0.0, 0.0, was_open=False) is False      # a closed gate stays closed
    assert gate_is_open(0, 5, 60.0, 0.0, 0.0, was_open=True) is True
    assert gate_decision(100, -1, 60.0, 0.0, 99.0) == (True, True)            # no ACK yet: open + timer restart
    assert gate_decision(40000, 10, 60.0, 0.0, 99.0) == (True, True)          # implausible gap: open + timer restart
    assert gate_decision(100, 90, 60.0, 0.0, 5.0) == (False, False)           # silent client


def test_channel_requires_an_event_loop_at_construction():
    """The callback thread has no loop of its own: a channel built without one must fail at construction, not lose stripes."""
    import pytest
    from selkies_b200.ws_video import WsVideoChannel

    async def send(_):
        pass
    with pytest.raises(RuntimeError):
        WsVideoChannel(send)                   # no running loop, none given

    async def inside():
        return WsVideoChannel(send)            # picks up the running loop
    ch = asyncio.run(inside())
    assert ch._loop is not None
