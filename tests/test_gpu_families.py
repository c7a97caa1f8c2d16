"""Tag families other than tag36h11, several families at once, and the reference's second detection path.

Reference: apriltag_utils.cu:10-27 (the families a caller may add), apriltag_gpu.cu:169-177 (min width_at_border and
border polarity over the families), gpu_detector_test.cu:122-157 (CpuAndGpuEqual: the classic CPU detector and the GPU
detector agree on id and on centre / corners within 0.5 px)."""
import numpy as np
import pytest

from parity import compare_all
from test_classic_detector import cpu_and_gpu_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    from ros_vision_b200 import detector
    detector.load_library()
    return detector


@pytest.mark.parametrize("family,ids", [("tag25h9", [0, 7, 34]), ("tag16h5", [1, 12, 29]), ("tag36h11", [3, 300, 586])])
def test_single_family_matches_oracle(D, oracle, family, ids):
    from ros_vision_b200 import synth
    for dec, side in ((2, (70, 130)), (1, (40, 90))):
        sc = synth.make_scene(640, 480, 500 + dec, 3, side_range=side, noise_sigma=3.0, family=family, ids=ids)
        det = D.GpuDetector(640, 480, "gray", quad_decimate=dec, keep_stages=True, families=[family])
        det.Detect(sc.gray)
        orc = oracle.detect(oracle.make_config(640, 480, "gray", dec, 0.0, families=[family]), sc.gray)
        got = compare_all(det, orc, 0, "gray")
        found = sorted(int(d["id"]) for d in got if int(d["hamming"]) == 0)
        assert found == sorted(ids)
        det.close()


def test_family_given_as_table_equals_builtin(D, oracle):
    """A caller-supplied family (what the C++ class builds from the caller's apriltag_family_t) against the built-in one."""
    from ros_vision_b200 import synth
    from ros_vision_b200.tag_families import FAMILIES
    sc = synth.make_scene(640, 480, 61, 3, side_range=(70, 130), noise_sigma=3.0, family="tag25h9", ids=[2, 9, 20])
    table = dict(FAMILIES["tag25h9"], name="mine")
    a = D.GpuDetector(640, 480, "gray", families=[table])
    b = D.GpuDetector(640, 480, "gray", families=["tag25h9"])
    a.Detect(sc.gray)
    b.Detect(sc.gray)
    assert np.array_equal(a.Detections(), b.Detections()) and len(a.Detections()) == 3
    a.close()
    b.close()


def test_several_families_at_once(D, oracle):
    """Two and three families on one detector: min_tag_width is the minimum over them (apriltag_gpu.cu:169-177), every
    quad is decoded against each, detections carry the index of their family."""
    from ros_vision_b200 import synth
    a = synth.make_scene(640, 480, 71, 2, side_range=(80, 120), noise_sigma=0.0, family="tag36h11", ids=[10, 11])
    b = synth.make_scene(640, 480, 72, 2, side_range=(80, 120), noise_sigma=0.0, family="tag16h5", ids=[4, 5])
    # left half from one scene, right half from the other
    img = a.gray.copy()
    img[:, 320:] = b.gray[:, 320:]
    img = np.clip(img.astype(np.int32) + np.random.default_rng(0).normal(0, 3, img.shape), 0, 255).astype(np.uint8)
    fams = ["tag36h11", "tag16h5", "tag25h9"]
    det = D.GpuDetector(640, 480, "gray", keep_stages=True, families=fams)
    det.Detect(img)
    orc = oracle.detect(oracle.make_config(640, 480, "gray", 2, 0.0, families=fams), img)
    got = compare_all(det, orc, 0, "gray")
    assert set(int(f) for f in got["family"]) <= {0, 1, 2}
    clean = [(int(d["family"]), int(d["id"])) for d in got if int(d["hamming"]) == 0]
    assert all(f in (0, 1) for f, _ in clean) and len(clean) >= 1
    det.close()


def test_invalid_family_sets_are_refused(D):
    from ros_vision_b200.tag_families import FAMILIES
    rev = dict(FAMILIES["tag16h5"], reversed_border=1, name="reversed16h5")
    with pytest.raises(D.B200TagError):     # mixed border polarities (apriltag_detect.cu:108)
        D.GpuDetector(640, 480, "gray", families=["tag36h11", rev])
    with pytest.raises(D.B200TagError):
        D.GpuDetector(640, 480, "gray", families=[dict(FAMILIES["tag16h5"], total_width=20)])
    with pytest.raises(ValueError):
        D.GpuDetector(640, 480, "gray", families=["tagNope"])
    det = D.GpuDetector(640, 480, "gray", families=[rev])   # a reversed-border family alone is fine
    det.Detect(np.full((480, 640), 128, np.uint8))
    assert len(det.Detections()) == 0
    det.close()


@pytest.mark.parametrize("name", ["ref_colorimage_crop", "ref_colorimage_notags_crop", "ref_grayimage_crop", "synthetic_cfg1"])
def test_cpu_and_gpu_equal(D, oracle, name):
    """gpu_detector_test.cu:122-157 with the CUDA engine as the GPU detector and oracle/classic_detector.c as the CPU
    detector: same ids, centre and corners within 0.5 px (the reference's tolerance)."""
    from helpers import load_golden
    meta, img = load_golden(name)
    w, h = meta["width"], meta["height"]
    det = D.GpuDetector(w, h, "gray")
    det.Detect(img)
    cpu, _ = oracle.classic_detect(oracle.make_config(w, h, "gray", 2, 0.0), img)
    cpu_and_gpu_equal(cpu, det.Detections())
    det.close()


def test_cpu_and_gpu_equal_on_bench_frames(D, oracle):
    from ros_vision_b200 import synth
    for cfgid, count in ((2, 4), (4, 2)):
        for i in range(count):
            frame, fmt, w, h, dec, sigma, sc = synth.config_frame(cfgid, i)
            det = D.GpuDetector(w, h, fmt, quad_decimate=dec)
            det.Detect(frame)
            gpu = det.Detections()
            cpu, _ = oracle.classic_detect(oracle.make_config(w, h, "gray", dec, sigma), sc.gray)
            both = sorted(set(int(d["id"]) for d in cpu) & set(int(d["id"]) for d in gpu))
            assert len(both) >= max(1, len(sc.tags) - 1)
            cpu_and_gpu_equal([d for d in cpu if int(d["id"]) in both], [d for d in gpu if int(d["id"]) in both])
            det.close()
