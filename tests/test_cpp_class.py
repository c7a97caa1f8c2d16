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


def test_class_library_exports_typed_debug_accessors(built):
    """apriltag_gpu.h:111-183: the thirteen accessors typed to the reference's packed records."""
    core, _ = built
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", core], text=True)
    for want in ["CopyUnionMarkerPairTo(frc971::apriltag::QuadBoundaryPoint*) const",
                 "CopyCompressedUnionMarkerPairTo(frc971::apriltag::QuadBoundaryPoint*) const", "CopySortedUnionMarkerPair() const",
                 "CopyExtents() const", "CopySelectedExtents() const", "CopySelectedBlobs() const", "CopySortedSelectedBlobs() const",
                 "CopyLineFitPoints() const", "CopyErrors() const", "CopyFilteredErrors() const", "CopyPeaks() const",
                 "NumCompressedPeaks() const", "CopyCompressedPeaks() const", "CopyFitQuads() const"]:
        assert "frc971::apriltag::GpuDetector::" + want in syms, want


def test_reference_record_layouts(tmp_path):
    """sizeof / bit positions of the reference's records (points.h:25-279, line_fit_filter.h:14-135)."""
    src = tmp_path / "layout.cc"
    src.write_text(r"""
#include <cstdio>
#include <cstddef>
#include "apriltags_cuda/reference_types.h"
using namespace frc971::apriltag;
int main() {
  QuadBoundaryPoint q; q.set_rep0(0x12345); q.set_rep1(0xabcde); q.set_base_xy(777, 333); q.set_dxy(3); q.set_black_to_white(true);
  IndexPoint ip(0x7ff, q.point_bits()); ip.set_theta(0xabcdef1);
  std::printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(QuadBoundaryPoint), sizeof(IndexPoint), sizeof(MinMaxExtents), sizeof(LineFitPoint),
              sizeof(LineFitMoments), sizeof(Peak), sizeof(FitQuad), offsetof(MinMaxExtents, pxgx_plus_pygy_sum));
  std::printf("%llx %u %u %d %d %llx %u %u\n", (unsigned long long)q.key, q.x(), q.y(), (int)q.gx(), (int)q.gy(), (unsigned long long)ip.key,
              ip.blob_index(), ip.theta());
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    assert out[0].split() == ["8", "8", "32", "48", "48", "12", "208", "24"]
    key = (0xabcde << 44) | (0x12345 << 24) | (777 << 14) | (333 << 4) | 8 | 3
    ikey = (0x7ff << 52) | (0xabcdef1 << 24) | (key & 0xffffff)
    assert out[1].split() == [f"{key:x}", str(2 * 777 - 1), str(2 * 333 + 1), "-1", "1", f"{ikey:x}", str(0x7ff), str(0xabcdef1)]


@pytest.mark.gpu
def test_typed_debug_accessors_against_oracle(built, oracle, tmp_path):
    """The typed accessors on the config-1 frame: counts, order invariants and order-independent sums against the oracle."""
    import numpy as np
    from ros_vision_b200 import synth
    frame, fmt, w, h, dec, sigma, sc = synth.config_frame(1)
    raw = tmp_path / "gray.raw"
    raw.write_bytes(frame.tobytes())
    exe = os.path.join(os.path.dirname(built[1]), "debug_accessors_test")
    r = subprocess.run([exe, str(raw), str(w), str(h)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(tok.split("=") for line in r.stdout.strip().splitlines() for tok in line.split())
    o = oracle.detect(oracle.make_config(w, h, fmt, dec, sigma), frame)
    sel = o.clusters[o.clusters["selected"] != 0]
    assert int(kv["points"]) == int(kv["dense_nonzero"]) == int(kv["count_sum"]) == len(o.points)
    assert int(kv["xsum"]) == int(o.points["x"].astype(np.int64).sum()) and int(kv["ysum"]) == int(o.points["y"].astype(np.int64).sum())
    assert int(kv["pairs"]) == int(kv["numquads"]) == len(o.clusters)
    assert int(kv["dot_sum"]) == int(o.clusters["pxgx_plus_pygy_sum"].sum())
    assert int(kv["selected_pairs"]) == len(sel)
    assert int(kv["selected_points"]) == int(kv["numselected"]) == int(kv["index_points"]) == int(kv["lfp"]) == len(o.spoints)
    assert int(kv["theta_sum"]) == int(o.spoints["theta"].astype(np.int64).sum()) and kv["same_sum"] == "1"
    last = np.cumsum(sel["count"].astype(np.int64)) - 1
    assert int(kv["w_last"]) == int(o.lfps["W"][last].sum())
    assert abs(float(kv["esum"]) - float(o.errs.sum())) <= 1e-6 * max(1.0, abs(float(o.errs.sum())))
    assert abs(float(kv["fsum"]) - float(o.filtered_errs.sum())) <= 1e-6 * max(1.0, abs(float(o.filtered_errs.sum())))
    assert int(kv["is_peak"]) == int(kv["compressed"]) == int(o.is_peak.sum())
    assert int(kv["fitquads"]) == int(kv["numfitquads"]) == len(o.fitquads)
    assert int(kv["valid"]) == int((o.fitquads["valid"] != 0).sum())
    assert int(kv["detections"]) == 4
