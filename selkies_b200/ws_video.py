"""WebSocket-mode consumer of the encoder callback: one queue, one sender task and an ACK-driven gate
per display.

What the reference does, spread over `queue_data_for_display` (selkies.py:3110-3135), `_video_chunk_sender` (:3026-3089),
the `CLIENT_FRAME_ACK` handler (:2225-2253) and `_run_frame_backpressure_logic` (:1196-1267):

* the native callback copies the stripe (10-byte header kept) and hands it to the event loop; a full queue drops the stripe;
* the sender stamps `(frame_id -> send time)` and sends, unless the gate is closed;
* the client acknowledges frame ids; an ACK yields an RTT sample (mean of the last 20) and resets the stall timer;
* every 0.5 s the gate is re-evaluated: closed when the client is more than 2 s of frames behind (frame ids are u16 and wrap;
  the RTT, when above 50 ms, is credited), or silent for 4 s; an implausible gap (> half the id range) or no ACK yet opens it.

`WsVideoChannel` packages that per display with the arithmetic in one pure function (`gate_is_open`) so it can be tested without
sockets.  The transport is any `async send(bytes)` callable.
"""
from __future__ import annotations

import asyncio
import time
from collections import OrderedDict, deque
from typing import Awaitable, Callable, Optional

ID_MODULUS = 65536                 # frame ids are u16 (selkies.py:10)
ALLOWED_LAG_MS = 2000.0            # selkies.py:7
RTT_CREDIT_ABOVE_MS = 50.0         # selkies.py:8
CHECK_PERIOD_S = 0.5               # selkies.py:9
STALL_AFTER_S = 4.0                # selkies.py:14
RTT_WINDOW = 20                    # selkies.py:15
SEND_STAMPS_KEPT = 1000            # selkies.py:16


def frames_behind(sent_id: int, acked_id: int) -> int:
    """How many frames the client is behind, on the u16 ring."""
    return (sent_id - acked_id) % ID_MODULUS


def gate_decision(sent_id: int, acked_id: int, fps: float, rtt_ms: float, silent_for_s: float):
    """One pass of selkies.py:1218-1260.  Returns (open, reset_stall_timer); open is True (send), False (hold) or None (leave the
    gate as it is: the server has not sent frame id 0 -> N yet, selkies.py:1239)."""
    if acked_id < 0:                                             # nothing acknowledged yet (:1221-1226)
        return True, True
    if abs(sent_id - acked_id) > (ID_MODULUS - 1) // 2:          # not a lag, a reset on one side (:1234-1237)
        return True, True
    if sent_id == 0:
        return None, False
    if silent_for_s > STALL_AFTER_S:
        return False, False
    fps = fps if fps > 0 else 60.0
    credit = rtt_ms / 1000.0 * fps if rtt_ms > RTT_CREDIT_ABOVE_MS else 0.0
    return frames_behind(sent_id, acked_id) - credit <= ALLOWED_LAG_MS / 1000.0 * fps, False


def gate_is_open(sent_id: int, acked_id: int, fps: float, rtt_ms: float, silent_for_s: float, was_open: bool = True) -> bool:
    """True = keep sending.  `acked_id` < 0: nothing acknowledged yet."""
    o, _ = gate_decision(sent_id, acked_id, fps, rtt_ms, silent_for_s)
    return was_open if o is None else o


class WsVideoChannel:
    def __init__(self, send: Callable[[bytes], Awaitable[None]], loop: Optional[asyncio.AbstractEventLoop] = None,
                 queue_depth: int = 120, fps: float = 60.0, jpeg: bool = False, clock=time.monotonic):
        # `on_stripe` runs on the encoder's native thread, which has no event loop of its own: the loop is fixed HERE (the
        # reference captures it the same way, selkies.py:3132 uses the server's loop), never looked up from the callback.
        if loop is None:
            try:
                loop = asyncio.get_running_loop()
            except RuntimeError:
                raise RuntimeError("WsVideoChannel needs an event loop: construct it inside the loop or pass loop=") from None
        self._send, self._loop, self._clock = send, loop, clock
        self.queue: asyncio.Queue = asyncio.Queue(maxsize=queue_depth)
        self.fps, self.jpeg = fps, jpeg
        self.open = True                          # the gate ("backpressure_enabled" in the reference's display state)
        self.last_sent_id, self.acked_id = 0, -1
        self.last_ack_at = clock()
        self._stamps: "OrderedDict[int, float]" = OrderedDict()
        self._rtts: deque = deque(maxlen=RTT_WINDOW)
        self.dropped = self.sent = self.bytes_sent = 0

    # -- producer side: called on the encoder's native output thread ------------------------------------
    def on_stripe(self, result_ptr, _user=None) -> None:
        """`StripeCallback` target.  The view dies with the call: copy once, then hop to the event loop."""
        if not result_ptr:
            return
        res = result_ptr.contents
        if res.size <= 0:
            return
        chunk = bytes(res.data[:res.size])
        if self.jpeg:
            chunk = b"\x03\x00" + chunk           # JPEG stripes carry a 2-byte type prefix (selkies.py:3118)
        self._loop.call_soon_threadsafe(self._offer, chunk, int(res.frame_id))

    def _offer(self, chunk: bytes, frame_id: int) -> None:
        try:
            self.queue.put_nowait((chunk, frame_id))
        except asyncio.QueueFull:
            self.dropped += 1

    # -- consumer side ------------------------------------------------------------------------------------
    async def run_sender(self) -> None:
        while True:
            chunk, frame_id = await self.queue.get()
            try:
                if self.open:
                    self._stamps[frame_id] = self._clock()
                    while len(self._stamps) > SEND_STAMPS_KEPT:
                        self._stamps.popitem(last=False)
                    self.last_sent_id = frame_id
                    await self._send(chunk)
                    self.sent += 1
                    self.bytes_sent += len(chunk)
            finally:
                self.queue.task_done()

    def on_ack(self, message: str) -> bool:
        """`CLIENT_FRAME_ACK <id>`; returns False for a malformed message."""
        parts = message.split()
        if len(parts) < 2 or parts[0] != "CLIENT_FRAME_ACK":
            return False
        try:
            fid = int(parts[-1])
        except ValueError:
            return False
        now = self._clock()
        self.acked_id, self.last_ack_at = fid, now
        t0 = self._stamps.pop(fid, None)
        if t0 is not None and now >= t0:
            self._rtts.append((now - t0) * 1000.0)
        return True

    @property
    def rtt_ms(self) -> float:
        return sum(self._rtts) / len(self._rtts) if self._rtts else 0.0

    def evaluate_gate(self, client_fps: float = 0.0) -> bool:
        o, reset = gate_decision(self.last_sent_id, self.acked_id, client_fps or self.fps, self.rtt_ms,
                                 self._clock() - self.last_ack_at)
        if reset:
            self.last_ack_at = self._clock()      # no ACK yet / implausible id gap: the stall timer restarts (selkies.py:1225, 1236)
        if o is not None:
            self.open = o
        return self.open

    async def run_gate(self, client_fps: Callable[[], float] = lambda: 0.0) -> None:
        try:
            while True:
                await asyncio.sleep(CHECK_PERIOD_S)
                self.evaluate_gate(client_fps())
        finally:
            self.open = True
