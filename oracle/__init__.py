"""CPU oracle for the video hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under selkies_b200/ does (tests/test_boundary.py enforces it).

PARITY UNPINNED: see the headers of csc_ref.c and h264_ref.c — the reference tree holds neither an
implementation nor golden vectors for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB):
                raise
        _lib = C.CDLL(_LIB)
    return _lib


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def csc_nv12(bgra: np.ndarray, dst_w: int = 0, dst_h: int = 0, coded_w: int = 0, coded_h: int = 0):
    """BGRA (H,W,4) uint8 -> (Y (coded_h,coded_w), UV (coded_h/2,coded_w)) per oracle/csc_ref.c."""
    assert bgra.dtype == np.uint8 and bgra.ndim == 3 and bgra.shape[2] == 4
    bgra = np.ascontiguousarray(bgra)
    sh, sw = bgra.shape[:2]
    dst_w = dst_w or sw
    dst_h = dst_h or sh
    coded_w = coded_w or dst_w
    coded_h = coded_h or dst_h
    y = np.empty((coded_h, coded_w), np.uint8)
    uv = np.empty((coded_h // 2, coded_w), np.uint8)
    rc = lib().b2v_ref_csc_nv12(_u8p(bgra), sw, sh, sw * 4, dst_w, dst_h, coded_w, coded_h, _u8p(y), _u8p(uv))
    if rc != 0:
        raise ValueError(f"b2v_ref_csc_nv12 rc={rc}")
    return y, uv
