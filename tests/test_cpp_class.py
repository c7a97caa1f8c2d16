"""The header-compatible C++ class frc971::apriltag::GpuDetector (include/apriltags_cuda/apriltag_gpu.h)."""
import os
import subprocess

import pytest

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from ros_vision_b200 import build
    build.build_native()
    return build.build_cpp_class(), build.build_cpp_test()


def test_class_library_exports_reference_api(built):
    """Same mangled names a caller compiled against the reference header would look up."""
    core, _ = built
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", core], text=True)
    for want in ["frc971::apriltag::GpuDetector::GpuDetector(unsigned long, unsigned long, apriltag_detector*, "
                 "frc971::apriltag::CameraMatrix, frc971::apriltag::DistCoeffs)",
                 "frc971::apriltag::GpuDetector::Detect(unsigned char const*)",
                 "frc971::apriltag::GpuDetector::DetectMjpg(unsigned char const*, unsigned long)",
                 "frc971::apriltag::GpuDetector::FitQuads() const",
                 "frc971::apriltag::GpuDetector::ReinitializeDetections()",
                 "frc971::apriltag::GpuDetector::CopyGrayTo(unsigned char*) const",
                 "frc971::apriltag::GpuDetector::CopyThresholdedTo(unsigned char*) const",
                 "frc971::apriltag::GpuDetector::CopyUnionMarkersTo(unsigned int*) const",
                 "frc971::apriltag::GpuDetector::AdjustCenter(float (*) [2]) const",
                 "frc971::apriltag::GpuDetector::UnDistort(double*, double*, frc971::apriltag::CameraMatrix const*, "
                 "frc971::apriltag::DistCoeffs const*)",
                 "frc971::apriltag::GpuDetector::~GpuDetector()"]:
        assert want in syms, want


@pytest.mark.gpu
@pytest.mark.parametrize("name,count,tag_id", [("ref_colorimage_crop", 1, 554), ("ref_colorimage_notags_crop", 0, None),
                                               ("ref_grayimage_crop", 1, 585)])
def test_gpu_detector_test_cc(built, tmp_path, name, count, tag_id):
    """gpu_detector_test.cu:84-102 (GpuDetectsAprilTag / GpuNoAprilTagDetections) through the C++ class."""
    _, exe = built
    meta, img = load_golden(name)
    raw = tmp_path / "gray.raw"
    raw.write_bytes(img.tobytes())
    args = [exe, str(raw), str(meta["width"]), str(meta["height"]), str(count)] + ([str(tag_id)] if tag_id is not None else [])
    mjpg = False
    if tag_id is not None:  # the same frame as a JPEG bitstream through GpuDetector::DetectMjpg
        try:
            import cv2
            ok, buf = cv2.imencode(".jpg", img.reshape(meta["height"], meta["width"]), [cv2.IMWRITE_JPEG_QUALITY, 95])
            (tmp_path / "frame.jpg").write_bytes(buf.tobytes())
            args.append(str(tmp_path / "frame.jpg"))
            mjpg = True
        except ImportError:
            pass
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"detections={count}" in r.stdout
    if mjpg:
        assert f"mjpg detections={count}" in r.stdout
