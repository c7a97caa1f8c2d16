"""Camera wire format: JPEG bitstreams -> nvJPEG luminance decode on the detector's stream ->
the gray pipeline.  JPEG decoders are not bit-identical to one another (IDCT rounding), so parity is stated against the
luminance plane the engine actually decoded: every stage behind it must match the oracle run on that plane bit for
bit, and the plane itself must be within 2 grey levels of OpenCV's (libjpeg) decode of the same bitstream."""
import ctypes as C

import numpy as np
import pytest

from helpers import match_corner_sets
from parity import compare_all

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


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
def test_mjpg_frames_match_oracle_on_decoded_luminance(D, oracle, w, h, colour, seed):
    from ros_vision_b200 import synth
    sc = synth.make_scene(w, h, seed, 4, side_range=(50, 140), noise_sigma=3.0)
    jpg = _encode(sc.gray, colour)
    det = D.GpuDetector(w, h, "gray", quad_decimate=2, keep_stages=True)
    det.DetectMjpg([jpg])
    assert det.mjpg_backend in ("gpu", "hardware", "hybrid", "default")
    luma = det.CopyGrayTo(0).reshape(h, w)
    ref = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_GRAYSCALE)
    diff = np.abs(luma.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= (2 if not colour else 3), diff.max()   # colour: OpenCV goes through BGR and back to gray
    orc = oracle.detect(oracle.make_config(w, h, "gray", 2, 0.0), np.ascontiguousarray(luma))
    got = compare_all(det, orc, 0, "gray")
    truth = {t.tag_id: t.corners for t in sc.tags}
    assert len(got) >= 1
    for d in got:
        assert int(d["id"]) in truth and match_corner_sets(d["p"], truth[int(d["id"])]) < 1.0
    det.close()


def test_mjpg_batch_equals_gray_batch(D):
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


def test_mjpg_rejects_bad_input(D):
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
