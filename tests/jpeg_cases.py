"""JPEG test streams (shared by the CPU oracle tests and the GPU decoder tests): every baseline variant a camera or
libjpeg can produce, encoded with OpenCV from a synthetic tag scene."""
import numpy as np


def strip_dht(jpeg: bytes) -> bytes:
    """Removes the DHT segments: what UVC / AVI MJPG frames look like (the standard tables K.3-K.6 are implied)."""
    out, pos = bytearray(jpeg[:2]), 2
    while True:
        m, L = jpeg[pos + 1], (jpeg[pos + 2] << 8) | jpeg[pos + 3]
        if m == 0xDA:
            out += jpeg[pos:]
            return bytes(out)
        if m != 0xC4:
            out += jpeg[pos:pos + 2 + L]
        pos += 2 + L


def dht_payload(jpeg: bytes) -> bytes:
    out, pos = b"", 2
    while True:
        m, L = jpeg[pos + 1], (jpeg[pos + 2] << 8) | jpeg[pos + 3]
        if m == 0xC4:
            out += jpeg[pos + 4:pos + 2 + L]
        if m == 0xDA:
            return out
        pos += 2 + L


def make_cases(w=648, h=488, seed=5):
    import cv2
    from ros_vision_b200 import synth
    sc = synth.make_scene(w, h, seed, 3, side_range=(60, 120), noise_sigma=3.0)
    bgr = synth.gray_to_bgr(sc.gray, np.random.default_rng(seed))
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    specs = [
        ("gray", sc.gray, []),
        ("444", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]),
        ("422", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
        ("420", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420]),
        ("440", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]),
        ("411", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]),
        ("422_rst7", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 7]),
        ("422_rst1", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 1]),
        ("420_optimised", bgr, [cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
        ("422_q50", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_QUALITY, 50]),
        ("422_q100", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_QUALITY, 100]),
        ("422_q10", bgr, [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_QUALITY, 10]),
    ]
    cases = {}
    for name, img, params in specs:
        ok, buf = cv2.imencode(".jpg", img, params)
        assert ok
        cases[name] = buf.tobytes()
    cases["422_no_dht"] = strip_dht(cases["422"])
    cases["gray_no_dht"] = strip_dht(cases["gray"])
    return sc, cases
