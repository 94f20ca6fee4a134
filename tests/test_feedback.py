"""Receiver-feedback bridge (REMB -> bitrate, PLI -> IDR); host logic only, no GPU."""
from selkies_b200.feedback import EncoderFeedback


class FakeCapture:
    def __init__(self):
        self.rates, self.idrs = [], 0

    def update_video_bitrate(self, kbps):
        self.rates.append(kbps)

    def request_idr_frame(self):
        self.idrs += 1


def test_remb_maps_to_kbps_with_headroom_and_limits():
    t = [0.0]
    cap = FakeCapture()
    fb = EncoderFeedback(cap, start_kbps=8000, clock=lambda: t[0])
    fb.target_bitrate = 4_000_000                     # the sender's REMB handler does exactly this assignment
    assert cap.rates == [3600] and fb.target_bitrate == 3_600_000
    fb.target_bitrate = 3_950_000                     # < 5 % move: ignored
    assert cap.rates == [3600]
    fb.target_bitrate = 8_000_000                     # increase inside the hold-off window: ignored
    assert cap.rates == [3600]
    t[0] = 1.0
    fb.target_bitrate = 8_000_000
    assert cap.rates == [3600, 7200]
    fb.target_bitrate = 2_000_000                     # decreases apply immediately
    assert cap.rates[-1] == 1800
    fb.target_bitrate = 10                            # clamped to the reference's 1..100 Mbit/s range
    assert cap.rates[-1] == 1000
    t[0] = 2.0
    fb.target_bitrate = 10**12
    assert cap.rates[-1] == 100000


def test_pli_bursts_collapse_into_one_idr():
    t = [0.0]
    cap = FakeCapture()
    fb = EncoderFeedback(cap, clock=lambda: t[0])
    assert fb.on_pli() and not fb.on_pli("peer", "viewer")
    t[0] = 0.6
    assert fb.on_pli()
    assert cap.idrs == 2 and fb.idr_requests == 2


def test_bridge_for_app_facade():
    class App:
        video_bitrate = 6000
        def __init__(self): self.calls = []
        def set_video_bitrate(self, k): self.calls.append(("rate", k))
        def send_idr(self): self.calls.append(("idr",))
    app = App()
    fb = EncoderFeedback.for_app(app)
    fb.target_bitrate = 3_000_000
    fb.on_pli()
    assert app.calls == [("rate", 2700), ("idr",)]
