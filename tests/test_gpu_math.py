"""The oracle's emulation of the device atan2f / hypotf must match the B200 bit for bit
(oracle/cuda_math_emul.h)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(op, a, b):
    from ros_vision_b200 import detector
    L = detector.load_library()
    out = np.zeros_like(a)
    rc = L.b200tag_debug_math(op, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), a.size)
    assert rc == 0
    return out


def _inputs():
    rng = np.random.default_rng(1)
    parts = [rng.normal(0, s, 200000).astype(np.float32) for s in (1e-3, 1.0, 50.0, 4000.0, 1e12)]
    ints = rng.integers(-255, 256, 200000).astype(np.float32)
    a = np.concatenate(parts + [ints, np.array([0.0, -0.0, 1.0, -1.0, 3.0, 4.0, np.inf, -np.inf, 1e-40, 5.0], np.float32)])
    b = np.concatenate([rng.permutation(p) for p in parts] + [rng.permutation(ints),
                       np.array([0.0, 0.0, -0.0, 0.0, 4.0, 3.0, np.inf, 1.0, 1e-42, 12.0], np.float32)])
    return a, b


@pytest.mark.parametrize("op,name", [(0, "orc_emul_atan2f"), (1, "orc_emul_hypotf")])
def test_emulation_is_bit_exact(oracle, op, name):
    a, b = _inputs()
    dev = _run(op, a, b)
    fn = getattr(oracle.lib(), name)
    # compare a strided subset element-wise through ctypes plus all special values
    idx = np.concatenate([np.arange(0, a.size, 37), np.arange(a.size - 10, a.size)])
    emu = np.array([fn(float(a[i]), float(b[i])) for i in idx], dtype=np.float32)
    assert np.array_equal(emu.view(np.uint32), dev[idx].view(np.uint32))
