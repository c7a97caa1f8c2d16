"""Camera wire format: JPEG bitstreams -> luminance decode on the detector's stream (the
engine's own kernel, csrc/kernels_jpeg.cu; nvJPEG for non-baseline streams) -> the gray pipeline.  JPEG decoders are
not bit-identical to one another (T.81 leaves the IDCT arithmetic open, T.83 bounds the error), so parity has two
parts: the decoded plane is within 1 grey level of the oracle's (oracle/jpeg_oracle.c, itself pinned within 1 level of
libjpeg-turbo by tests/test_jpeg_oracle.py), and every stage behind it matches the AprilTag oracle run on that plane
bit for bit."""
import ctypes as C

import numpy as np
import pytest

from helpers import match_corner_sets
from parity import compare_all

from jpeg_cases import make_cases

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(params=["native", "nvjpeg"])
def decoder(request, monkeypatch):
    if request.param == "nvjpeg":
        monkeypatch.setenv("B200TAG_MJPG_DECODER", "nvjpeg")
    else:
        monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    return request.param


@pytest.fixture(scope="module")
def D():
    from ros_vision_b200 import detector
    detector.load_library()
    return detector


def _encode(gray, colour, quality=92):
    from ros_vision_b200 import synth
    img = synth.gray_to_bgr(gray, np.random.default_rng(3)) if colour else gray
    params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    if colour:  # what UVC cameras send: YCbCr 4:2:2
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]
    ok, buf = cv2.imencode(".jpg", img, params)
    assert ok
    return buf.tobytes()


@pytest.mark.parametrize("w,h,colour,seed", [(1280, 800, True, 41), (1280, 800, False, 42), (640, 480, True, 43), (328, 248, True, 44)])
def test_mjpg_frames_match_oracle_on_decoded_luminance(D, oracle, decoder, w, h, colour, seed):
    from ros_vision_b200 import synth
    sc = synth.make_scene(w, h, seed, 4, side_range=(50, 140), noise_sigma=3.0)
    jpg = _encode(sc.gray, colour)
    det = D.GpuDetector(w, h, "gray", quad_decimate=2, keep_stages=True)
    det.DetectMjpg([jpg])
    assert det.mjpg_backend == "native" if decoder == "native" else det.mjpg_backend in ("gpu", "hardware", "hybrid", "default")
    luma = det.CopyGrayTo(0).reshape(h, w)
    from oracle import pyjpeg
    ref = pyjpeg.decode_luma(jpg)
    diff = np.abs(luma.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= (1 if decoder == "native" else 2), diff.max()
    orc = oracle.detect(oracle.make_config(w, h, "gray", 2, 0.0), np.ascontiguousarray(luma))
    got = compare_all(det, orc, 0, "gray")
    truth = {t.tag_id: t.corners for t in sc.tags}
    assert len(got) >= 1
    for d in got:
        assert int(d["id"]) in truth and match_corner_sets(d["p"], truth[int(d["id"])]) < 1.0
    det.close()


def test_mjpg_batch_equals_gray_batch(D, decoder):
    """A batch of JPEG frames gives exactly what the same detector gives for the decoded planes passed as gray frames."""
    from ros_vision_b200 import synth
    w, h, n = 1280, 800, 6
    scenes = [synth.make_scene(w, h, 500 + i, 3 + i % 3, side_range=(60, 200), noise_sigma=3.0) for i in range(n)]
    jpgs = [_encode(s.gray, True, quality=85 + i) for i, s in enumerate(scenes)]
    det = D.GpuDetector(w, h, "gray", quad_decimate=2, keep_stages=True, max_batch=n)
    det.DetectMjpg(jpgs)
    via_jpeg = [det.Detections(f).copy() for f in range(n)]
    planes = [det.CopyGrayTo(f).reshape(h, w).copy() for f in range(n)]
    det.DetectBatch(planes)
    for f in range(n):
        again = det.Detections(f)
        assert len(again) == len(via_jpeg[f]) >= 1
        assert np.array_equal(again["id"], via_jpeg[f]["id"]) and np.array_equal(again["p"], via_jpeg[f]["p"])
        assert {int(t.tag_id) for t in scenes[f].tags} >= {int(i) for i in again["id"]}
    # a second batch of another size re-initialises the batched decoder
    det.DetectMjpg(jpgs[:2])
    assert np.array_equal(det.Detections(1)["p"], via_jpeg[1]["p"])
    det.close()


def test_native_decoder_every_baseline_variant(D, monkeypatch):
    """Gray / 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 / 4:1:1, restart intervals, optimised and implied Huffman tables, quality 10
    to 100, a frame size that is not a multiple of the MCU: the kernel's plane against the oracle's, one batch."""
    from oracle import pyjpeg
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    sc, streams = make_cases()
    h, w = sc.gray.shape
    names = sorted(streams)
    det = D.GpuDetector(w, h, "gray", quad_decimate=2, keep_stages=True, max_batch=len(names))
    det.DetectMjpg([streams[n] for n in names])
    assert det.mjpg_backend == "native"
    # every stream takes the parallel kernels (a frame whose synchronisation is not proven within the fixed number of
    # rounds, or whose restart markers do not add up, would fall back to the sequential kernel: none of these does)
    assert det.MjpgParallelFrames() == len(names)
    planes = [det.CopyGrayTo(f).reshape(h, w).copy() for f in range(len(names))]
    monkeypatch.setenv("B200TAG_MJPG_DECODER", "sequential")
    det.DetectMjpg([streams[n] for n in names])
    assert det.MjpgParallelFrames() == 0
    monkeypatch.delenv("B200TAG_MJPG_DECODER")
    exact = 0
    for f, name in enumerate(names):
        ref = pyjpeg.decode_luma(streams[name])
        got = planes[f]
        assert np.array_equal(got, det.CopyGrayTo(f).reshape(h, w)), name   # parallel and sequential kernels: same bits
        assert np.array_equal(got, D.jpeg_model_decode(streams[name], w, h)[0]), name   # and the host model
        diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1, (name, diff.max())
        exact += int((diff == 0).mean() > 0.999)   # float vs double IDCT: ties at .5 are the only differences
        if "q10" not in name:
            assert {int(i) for i in det.Detections(f)["id"]} == {int(t.tag_id) for t in sc.tags}, name
    assert exact == len(names)
    det.close()


def test_native_decoder_partial_blocks_at_the_edges(D, monkeypatch):
    """A frame size that is a multiple of neither 8 nor 16 (the detector itself needs multiples of 4 at quad_decimate 1):
    clipped blocks, pixel rows that are not 8-byte aligned."""
    from oracle import pyjpeg
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    w, h = 364, 252
    sc, streams = make_cases(w=w, h=h, seed=8)
    names = ["gray", "444", "422", "420", "411", "422_rst7", "420_optimised"]
    det = D.GpuDetector(w, h, "gray", quad_decimate=1, keep_stages=True, max_batch=len(names))
    det.DetectMjpg([streams[n] for n in names], allow_overflow=True)
    assert det.MjpgParallelFrames() == len(names)
    for f, name in enumerate(names):
        got = det.CopyGrayTo(f).reshape(h, w)
        assert np.abs(got.astype(np.int32) - pyjpeg.decode_luma(streams[name]).astype(np.int32)).max() <= 1, name
        assert np.array_equal(got, D.jpeg_model_decode(streams[name], w, h)[0]), name
    det.close()


def test_native_decoder_at_the_largest_baseline_size(D, monkeypatch):
    """BASELINE config 5's frame size (3840x2160, 100 tags + clutter) as one 4:2:0 JPEG: 130 k blocks, 20 k subsequences."""
    from oracle import pyjpeg
    from ros_vision_b200 import synth
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    _, _, w, h, dec, sigma, sc = synth.config_frame(5, 0)
    ok, buf = cv2.imencode(".jpg", synth.gray_to_bgr(sc.gray, np.random.default_rng(5)), [cv2.IMWRITE_JPEG_QUALITY, 85])
    jpg = buf.tobytes()
    det = D.GpuDetector(w, h, "gray", quad_decimate=dec, quad_sigma=sigma, keep_stages=True)
    det.DetectMjpg([jpg], allow_overflow=True)
    assert det.MjpgParallelFrames() == 1
    got = det.CopyGrayTo(0).reshape(h, w)
    ref = pyjpeg.decode_luma(jpg)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-4
    found = {int(i) for i in det.Detections(0)["id"]}
    assert len(found & {int(t.tag_id) for t in sc.tags}) >= 0.9 * len(sc.tags)
    det.close()


def test_mjpg_random_streams(D, monkeypatch):
    """Random scenes, sizes, qualities, samplings, restart intervals and Huffman optimisation, in random batches: every
    plane within 1 level of the JPEG oracle and identical to the host model, every frame on the parallel kernels.
    B200TAG_SWEEP_SEED / B200TAG_SWEEP_COUNT run a longer sweep with other draws."""
    import os
    from oracle import pyjpeg
    from ros_vision_b200 import synth
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    seed, count = int(os.environ.get("B200TAG_SWEEP_SEED", "11")), int(os.environ.get("B200TAG_SWEEP_COUNT", "12"))
    rng = np.random.default_rng(seed)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    samplings = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
    done = 0
    while done < count:
        w, h = int(rng.integers(12, 80)) * 8, int(rng.integers(10, 60)) * 8
        n = int(rng.integers(1, 6))
        det = D.GpuDetector(w, h, "gray", quad_decimate=2, keep_stages=True, max_batch=n)
        jpgs = []
        for i in range(n):
            sc = synth.make_scene(w, h, int(rng.integers(1, 1 << 30)), int(rng.integers(0, 4)), side_range=(24.0, max(30.0, min(w, h) / 2.5)),
                                  noise_sigma=float(rng.uniform(0, 8)), clutter=bool(rng.integers(0, 2)))
            colour = bool(rng.integers(0, 4))
            img = synth.gray_to_bgr(sc.gray, np.random.default_rng(i)) if colour else sc.gray
            params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(5, 101))]
            if colour:
                params += [S, int(rng.choice(samplings))]
            if rng.integers(0, 3) == 0:
                params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 200))]
            if rng.integers(0, 3) == 0:
                params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
            ok, buf = cv2.imencode(".jpg", img, params)
            assert ok
            jpgs.append(buf.tobytes())
        det.DetectMjpg(jpgs, allow_overflow=True)
        assert det.MjpgParallelFrames() == n
        for f, jpg in enumerate(jpgs):
            got = det.CopyGrayTo(f).reshape(h, w)
            assert np.abs(got.astype(np.int32) - pyjpeg.decode_luma(jpg).astype(np.int32)).max() <= 1
            assert np.array_equal(got, D.jpeg_model_decode(jpg, w, h)[0])
        det.close()
        done += n


def test_non_baseline_streams_go_through_nvjpeg(D, monkeypatch):
    from ros_vision_b200 import synth
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    w, h = 640, 480
    sc = synth.make_scene(w, h, 21, 3, side_range=(60, 120), noise_sigma=2.0)
    ok, prog = cv2.imencode(".jpg", sc.gray, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    det = D.GpuDetector(w, h, "gray", quad_decimate=2)
    det.DetectMjpg([prog.tobytes()])
    assert det.mjpg_backend != "native"
    assert {int(i) for i in det.Detections(0)["id"]} == {int(t.tag_id) for t in sc.tags}
    det.DetectMjpg([_encode(sc.gray, True)])
    assert det.mjpg_backend == "native"
    assert {int(i) for i in det.Detections(0)["id"]} == {int(t.tag_id) for t in sc.tags}
    det.close()


def test_corrupt_entropy_data_does_not_hang_or_crash(D, monkeypatch):
    """Bit errors inside the entropy-coded segment: the frame decodes to garbage, the call returns, the next frame is fine."""
    from ros_vision_b200 import synth
    monkeypatch.delenv("B200TAG_MJPG_DECODER", raising=False)
    w, h = 640, 480
    sc = synth.make_scene(w, h, 22, 3, side_range=(60, 120), noise_sigma=2.0)
    good = _encode(sc.gray, True)
    rng = np.random.default_rng(1)
    bad = bytearray(good)
    for i in rng.integers(1000, len(bad) - 2, size=200):
        bad[i] = int(rng.integers(0, 256))
    det = D.GpuDetector(w, h, "gray", quad_decimate=2, max_batch=3)
    det.DetectMjpg([bytes(bad), good, good[:len(good) // 2]], allow_overflow=True)
    assert {int(i) for i in det.Detections(1)["id"]} == {int(t.tag_id) for t in sc.tags}
    # the cut-off frame is flagged so that the caller can drop it; complete frames are not
    assert det.FrameInfo(2).status & D.ST_JPEG_TRUNCATED
    assert not det.FrameInfo(1).status & D.ST_JPEG_TRUNCATED and not det.FrameInfo(0).status & D.ST_JPEG_TRUNCATED
    det.DetectMjpg([good])
    assert det.FrameInfo(0).status == 0
    det.close()


def test_mjpg_rejects_bad_input(D, decoder):
    from ros_vision_b200 import synth
    w, h = 640, 480
    sc = synth.make_scene(w, h, 9, 2, side_range=(60, 120))
    good = _encode(sc.gray, True)
    det = D.GpuDetector(w, h, "gray", quad_decimate=2)
    with pytest.raises(D.B200TagError, match="JPEG"):
        det.DetectMjpg([b"\x00" * 4096])
    with pytest.raises(D.B200TagError, match="320x240"):
        det.DetectMjpg([_encode(sc.gray[:240, :320].copy(), False)])
    det.DetectMjpg([good])  # still usable afterwards
    assert len(det.Detections(0)) >= 1
    det.close()
    yuyv = D.GpuDetector(w, h, "yuyv", quad_decimate=2)
    with pytest.raises(D.B200TagError, match="GRAY8"):
        yuyv.DetectMjpg([good])
    yuyv.close()
