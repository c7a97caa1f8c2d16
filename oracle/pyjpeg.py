"""TEST INFRASTRUCTURE: ctypes binding of oracle/libjpeg_oracle.so (jpeg_oracle.c, baseline JPEG -> luminance plane)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Info(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int), ("hs", C.c_int * 4), ("vs", C.c_int * 4),
                ("tq", C.c_int * 4), ("cid", C.c_int * 4), ("restart_interval", C.c_int)]


def load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libjpeg_oracle.so")
        src = os.path.join(_HERE, "jpeg_oracle.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libjpeg_oracle.so"])
        _lib = C.CDLL(so)
        _lib.jpeg_oracle_decode_luma.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Info)]
        _lib.jpeg_oracle_standard_dht.argtypes = [C.c_void_p, C.c_size_t]
    return _lib


def info(jpeg: bytes) -> Info:
    fi = Info()
    buf = np.frombuffer(jpeg, dtype=np.uint8)
    rc = load().jpeg_oracle_decode_luma(buf.ctypes.data, buf.size, None, 0, C.byref(fi))
    if rc:
        raise ValueError(f"jpeg_oracle_decode_luma: {rc}")
    return fi


def decode_luma(jpeg: bytes) -> np.ndarray:
    fi = info(jpeg)
    out = np.zeros((fi.height, fi.width), dtype=np.uint8)
    buf = np.frombuffer(jpeg, dtype=np.uint8)
    rc = load().jpeg_oracle_decode_luma(buf.ctypes.data, buf.size, out.ctypes.data, out.size, None)
    if rc:
        raise ValueError(f"jpeg_oracle_decode_luma: {rc}")
    return out


def standard_dht() -> bytes:
    out = np.zeros(1024, dtype=np.uint8)
    n = load().jpeg_oracle_standard_dht(out.ctypes.data, out.size)
    return out[:n].tobytes()
