"""The pip-`apriltag`-shaped Python surface that the reference's offline calibration tool
uses (extrinsic_calibration/solver.py:181-200), on top of the C ABI."""
import numpy as np
import pytest

from helpers import match_corner_sets

pytestmark = pytest.mark.gpu


def test_solver_style_usage():
    import ros_vision_b200.apriltag_api as apriltag
    from ros_vision_b200 import synth
    w, h = 1280, 800
    fx = fy = 900.0
    cx, cy = w / 2, h / 2
    sc = synth.make_scene(w, h, 123, 4, side_range=(120, 260), noise_sigma=3.0)
    options = apriltag.DetectorOptions(families="tag36h11")          # solver.py:181
    detector = apriltag.Detector(options)                            # :182
    detection_list = detector.detect(sc.gray)                        # :189
    truth = {t.tag_id: t.corners for t in sc.tags}
    assert sorted(d.tag_id for d in detection_list) == sorted(truth)
    for det in detection_list:
        assert det.corners.shape == (4, 2) and det.homography.shape == (3, 3) and det.center.shape == (2,)
        assert match_corner_sets(det.corners, truth[det.tag_id]) < 1.0
        solvedpose = detector.detection_pose(det, (fx, fy, cx, cy), tag_size=0.1651)   # :198
        pose = solvedpose[0]
        translation = pose[:3, 3]                                    # :199
        assert pose.shape == (4, 4) and translation[2] > 0.1
        # the pose reprojects the tag's corners onto the detected ones (loosely: the scene generator gives every tag
        # its own little pinhole model, so no single camera explains all four corners exactly)
        s = 0.1651 / 2
        obj = np.array([[-s, s, 0, 1], [s, s, 0, 1], [s, -s, 0, 1], [-s, -s, 0, 1]]).T
        cam = (pose @ obj)[:3]
        px = np.stack([fx * cam[0] / cam[2] + cx, fy * cam[1] / cam[2] + cy], axis=1)
        assert np.abs(px - det.corners).max() < 8.0
    # a second image size through the same Detector object
    sc2 = synth.make_scene(640, 480, 124, 2, side_range=(80, 160), noise_sigma=2.0)
    assert sorted(d.tag_id for d in detector.detect(sc2.gray)) == sorted(t.tag_id for t in sc2.tags)
    detector.close()


def test_unsupported_options_are_rejected():
    import ros_vision_b200.apriltag_api as apriltag
    with pytest.raises(ValueError):
        apriltag.Detector(apriltag.DetectorOptions(families="tag25h9"))
    with pytest.raises(ValueError):
        apriltag.Detector(apriltag.DetectorOptions(quad_decimate=1.5))
