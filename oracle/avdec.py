"""H.264 decode oracle: libavcodec's native `h264` decoder (the copy bundled with opencv-python-headless)
driven through ctypes.  TEST INFRASTRUCTURE ONLY.

This is the independent conformance check for the encoder: every access unit must
decode without error and the decoded planes must equal the encoder's own reconstruction bit-for-bit.
It is not the reference's encoder (libx264 is absent from this image) — it is a third-party decoder.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_av = None


class _AVPacket(C.Structure):       # prefix of struct AVPacket (libavcodec 62)
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int)]


class _AVFrame(C.Structure):        # prefix of struct AVFrame (libavutil 60)
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


def _load():
    global _av
    if _av is not None:
        return _av
    import cv2  # noqa: F401  (resolves the wheel's bundled libav* dependencies first)
    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    avutil = C.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0])
    avcodec = C.CDLL(glob.glob(os.path.join(d, "libavcodec-*.so*"))[0])
    avcodec.avcodec_find_decoder.restype = C.c_void_p
    avcodec.avcodec_find_decoder.argtypes = [C.c_int]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avcodec.av_packet_alloc.restype = C.POINTER(_AVPacket)
    avcodec.av_new_packet.argtypes = [C.POINTER(_AVPacket), C.c_int]
    avcodec.av_packet_unref.argtypes = [C.POINTER(_AVPacket)]
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    avutil.av_frame_alloc.restype = C.POINTER(_AVFrame)
    avutil.av_frame_unref.argtypes = [C.POINTER(_AVFrame)]
    avutil.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    avutil.av_log_set_level.argtypes = [C.c_int]
    _av = (avcodec, avutil)
    return _av


class DecodeError(RuntimeError):
    pass


class H264Decoder:
    """Feed Annex-B access units, get (Y, U, V) uint8 planes back (yuv420p, visible size)."""

    def __init__(self, quiet: bool = False):
        avcodec, avutil = _load()
        self.avcodec, self.avutil = avcodec, avutil
        avutil.av_log_set_level(8 if quiet else 16)            # AV_LOG_FATAL / AV_LOG_ERROR
        dec = avcodec.avcodec_find_decoder(27)                 # AV_CODEC_ID_H264
        if not dec:
            raise DecodeError("libavcodec has no h264 decoder")
        self.ctx = C.c_void_p(avcodec.avcodec_alloc_context3(dec))
        avutil.av_opt_set(self.ctx, b"flags", b"+low_delay", 0)
        avutil.av_opt_set(self.ctx, b"err_detect", b"+explode+bitstream+buffer+crccheck", 0)
        avutil.av_opt_set(self.ctx, b"threads", b"1", 0)
        if avcodec.avcodec_open2(self.ctx, dec, None) < 0:
            raise DecodeError("avcodec_open2 failed")
        self.pkt = avcodec.av_packet_alloc()
        self.frm = avutil.av_frame_alloc()

    def _drain(self, out):
        while True:
            rc = self.avcodec.avcodec_receive_frame(self.ctx, self.frm)
            if rc < 0:
                return rc
            f = self.frm.contents
            if f.format not in (0, 12):                         # yuv420p / yuvj420p
                raise DecodeError(f"unexpected pixel format {f.format}")
            planes = []
            for i, (w, h) in enumerate(((f.width, f.height), (f.width // 2, f.height // 2), (f.width // 2, f.height // 2))):
                ls = f.linesize[i]
                buf = (C.c_ubyte * (ls * h)).from_address(f.data[i])
                planes.append(np.frombuffer(buf, np.uint8).reshape(h, ls)[:, :w].copy())
            out.append(tuple(planes))
            self.avutil.av_frame_unref(self.frm)

    def decode(self, au: bytes):
        """Send one access unit; returns the list of frames that became available."""
        out = []
        if self.avcodec.av_new_packet(self.pkt, len(au)) < 0:
            raise DecodeError("av_new_packet failed")
        C.memmove(self.pkt.contents.data, au, len(au))
        rc = self.avcodec.avcodec_send_packet(self.ctx, self.pkt)
        self.avcodec.av_packet_unref(self.pkt)
        if rc < 0:
            raise DecodeError(f"avcodec_send_packet failed: {rc}")
        self._drain(out)
        return out

    def flush(self):
        out = []
        self.avcodec.avcodec_send_packet(self.ctx, None)
        self._drain(out)
        return out

    def close(self):
        if self.ctx:
            self.avcodec.avcodec_free_context(C.byref(self.ctx))
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def decode_stream(aus, quiet: bool = False):
    """Decode a list of access units; returns a list of (Y,U,V) in output order."""
    frames = []
    with H264Decoder(quiet=quiet) as d:
        for au in aus:
            frames += d.decode(bytes(au))
        frames += d.flush()
    return frames


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))
