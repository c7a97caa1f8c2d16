"""Pins oracle/jpeg_oracle.c (the CPU restatement of ITU-T T.81 baseline decoding, luminance only) against libjpeg-turbo
as shipped inside this image's OpenCV -- the decoder the reference's camera node uses through cv::VideoCapture
(camera_publisher.cpp:198,336) -- and checks the product's host-side header parser against both.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from jpeg_cases import dht_payload, make_cases, strip_dht

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def cases():
    return make_cases()


@pytest.fixture(scope="module")
def pyjpeg():
    from oracle import pyjpeg
    pyjpeg.load()
    return pyjpeg


def test_oracle_within_one_level_of_libjpeg(cases, pyjpeg):
    sc, streams = cases
    for name, jpg in streams.items():
        ref = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_GRAYSCALE)
        got = pyjpeg.decode_luma(jpg)
        assert got.shape == ref.shape == sc.gray.shape, name
        assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1, name


def test_streams_without_dht_use_the_standard_tables(cases, pyjpeg):
    _, streams = cases
    assert pyjpeg.standard_dht() == dht_payload(streams["422"])  # libjpeg's default tables are K.3 - K.6
    assert np.array_equal(pyjpeg.decode_luma(streams["422_no_dht"]), pyjpeg.decode_luma(streams["422"]))


def test_oracle_rejects_what_it_does_not_decode(cases, pyjpeg):
    _, streams = cases
    ok, prog = cv2.imencode(".jpg", np.zeros((64, 64, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError, match="-2"):
        pyjpeg.decode_luma(prog.tobytes())
    with pytest.raises(ValueError, match="-1"):
        pyjpeg.decode_luma(b"\x00" * 100)
    with pytest.raises(ValueError, match="-1"):
        pyjpeg.decode_luma(streams["422"][:300])


def _probe(lib, jpg):
    info = (C.c_int32 * 8)()
    dht = np.zeros(2048, np.uint8)
    n = C.c_size_t(0)
    buf = np.frombuffer(jpg, np.uint8)
    rc = lib.b200tag_jpeg_probe(buf.ctypes.data_as(C.c_void_p), buf.size, info, dht.ctypes.data_as(C.c_void_p), dht.size, C.byref(n))
    return rc, list(info), dht[:n.value].tobytes()


def test_host_parser_agrees_with_oracle(cases, pyjpeg):
    from ros_vision_b200 import build, detector
    build.build_native()
    lib = detector.load_library()
    _, streams = cases
    for name, jpg in streams.items():
        rc, info, dht = _probe(lib, jpg)
        fi = pyjpeg.info(jpg)
        assert rc == 0, name
        hmax, vmax = (fi.hs[0], fi.vs[0]) if fi.ncomp == 3 else (1, 1)
        nblocks = sum(fi.hs[c] * fi.vs[c] for c in range(fi.ncomp)) if fi.ncomp == 3 else 1
        assert info[:6] == [fi.width, fi.height, nblocks, hmax, vmax, fi.restart_interval], name
        assert jpg[info[6] - 3:info[6]] == bytes([0, 63, 0])  # Ss, Se, Ah/Al close the SOS header
        assert info[7] == -(-fi.width // (8 * hmax)) * -(-fi.height // (8 * vmax))
    # the tables the parser assumes for a stream without DHT are libjpeg's defaults (= T.81 K.3 - K.6)
    _, _, implied = _probe(lib, streams["422_no_dht"])
    _, _, explicit = _probe(lib, streams["422"])
    assert implied == explicit and len(implied) > 400
    ok, prog = cv2.imencode(".jpg", np.zeros((64, 64, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert _probe(lib, prog.tobytes())[0] == 1
    assert _probe(lib, b"\x12" * 64)[0] < 0
    assert _probe(lib, streams["422"][:200])[0] < 0


def test_parallel_scheme_host_model(cases, pyjpeg):
    """The self-synchronising parallel decode (subsequences, fixed-point rounds, block numbering, DC scan, float IDCT) as
    the kernels do it, run by the host model (restart-marker streams: one interval per thread instead): within 1 level
    of the oracle on every stream."""
    from ros_vision_b200 import build, detector
    build.build_native()
    sc, streams = cases
    h, w = sc.gray.shape
    for name, jpg in streams.items():
        got, rounds = detector.jpeg_model_decode(jpg, w, h)
        ref = pyjpeg.decode_luma(jpg)
        diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-4, name
        assert (rounds == 0) if "rst" in name else (1 <= rounds <= 200), (name, rounds)   # restart intervals need no rounds


def test_frame_size_not_a_multiple_of_the_block(pyjpeg):
    """Partial blocks / MCUs at the right and bottom edges (T.81 A.2.4), for oracle and host model alike."""
    from ros_vision_b200 import build, detector
    build.build_native()
    sc, streams = make_cases(w=362, h=251, seed=8)
    for name in ("gray", "444", "422", "420", "411", "422_rst7", "420_optimised"):
        jpg = streams[name]
        ref = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_GRAYSCALE)
        got = pyjpeg.decode_luma(jpg)
        assert got.shape == (251, 362) and np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1, name
        model, _ = detector.jpeg_model_decode(jpg, 362, 251)
        assert np.abs(model.astype(np.int32) - got.astype(np.int32)).max() <= 1, name


def test_corrupt_streams_do_not_crash_parser_or_model(cases):
    """Bit errors, truncation and garbage: the header parser and the host model of the decode kernels (the same
    entropy-decoding core the kernels compile) return -- a plane, or an error code -- and never read out of bounds."""
    import ctypes as C
    from ros_vision_b200 import build, detector
    build.build_native()
    lib = detector.load_library()
    sc, streams = cases
    h, w = sc.gray.shape
    rng = np.random.default_rng(123)
    out = np.zeros((h, w), np.uint8)
    rounds = C.c_int32(0)
    outcomes = {"decoded": 0, "rejected": 0}
    for name in ("422", "420_optimised", "422_rst7", "gray_no_dht"):
        base = bytearray(streams[name])
        for trial in range(60):
            bad = bytearray(base)
            kind = trial % 3
            if kind == 0:      # bit errors in the entropy-coded data
                for i in rng.integers(700, len(bad) - 2, size=int(rng.integers(1, 50))):
                    bad[i] = int(rng.integers(0, 256))
            elif kind == 1:    # truncation
                del bad[int(rng.integers(100, len(bad))):]
            else:              # errors in the headers
                for i in rng.integers(2, 650, size=int(rng.integers(1, 6))):
                    bad[i] = int(rng.integers(0, 256))
            buf = np.frombuffer(bytes(bad), np.uint8)
            rc = lib.b200tag_debug_jpeg_model(buf.ctypes.data_as(C.c_void_p), buf.size, out.ctypes.data_as(C.c_void_p), out.size, C.byref(rounds))
            assert rc in (0, 1, -1, -2), rc
            outcomes["decoded" if rc == 0 else "rejected"] += 1
    assert outcomes["decoded"] > 40 and outcomes["rejected"] > 20, outcomes
