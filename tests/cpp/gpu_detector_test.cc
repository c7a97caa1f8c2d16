// C++ mirror of the reference's src/apriltags_cuda/test/gpu_detector_test.cu (GpuDetectsAprilTag,
// GpuNoAprilTagDetections) against the header-compatible GpuDetector class, without gtest / OpenCV:
//   gpu_detector_test <gray.raw> <width> <height> <expected_count> [expected_id [frame.jpg]]
// The raw file is the luma plane of a golden fixture; it is packed to YUYV like the test's cvtColor.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "apriltags_cuda/apriltag_gpu.h"

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const int width = std::atoi(argv[2]), height = std::atoi(argv[3]), expected = std::atoi(argv[4]);
  std::vector<uint8_t> gray(static_cast<size_t>(width) * height);
  FILE *f = std::fopen(argv[1], "rb");
  if (!f || std::fread(gray.data(), 1, gray.size(), f) != gray.size()) return 3;
  std::fclose(f);
  std::vector<uint8_t> yuyv(gray.size() * 2, 128);
  for (size_t i = 0; i < gray.size(); i++) yuyv[2 * i] = gray[i];

  // gpu_detector_test.cu:49-73
  apriltag_family_t *tf = tag36h11_create();
  apriltag_detector_t *td = apriltag_detector_create();
  apriltag_detector_add_family(td, tf);
  td->quad_decimate = 2.0;
  td->quad_sigma = 0.0;
  td->nthreads = 1;
  td->debug = false;
  td->refine_edges = true;
  frc971::apriltag::CameraMatrix cam{905.495617, 609.916016, 907.909470, 352.682645};
  frc971::apriltag::DistCoeffs dist{0.059238, -0.075154, -0.003801, 0.001113, 0.0};
  int rc = 0;
  {
    frc971::apriltag::GpuDetector detector(width, height, td, cam, dist);
    detector.Detect(yuyv.data());
    const zarray_t *detections = detector.Detections();
    std::printf("detections=%d quads=%zu\n", zarray_size(detections), detector.FitQuads().size());
    if (zarray_size(detections) != expected) rc = 1;
    for (int i = 0; i < zarray_size(detections); i++) {
      apriltag_detection_t *det;
      zarray_get(detections, i, &det);
      std::printf("id=%d hamming=%d margin=%.3f c=(%.3f,%.3f) H22=%.3f\n", det->id, det->hamming, det->decision_margin, det->c[0],
                  det->c[1], matd_get(det->H, 2, 2));
      if (argc > 5 && det->id != std::atoi(argv[5])) rc = 1;
      // the node's next step, apriltags_cuda_detector.cu:425-462: pose of the tag in the camera frame
      apriltag_detection_info_t info{det, 0.1651, cam.fx, cam.fy, cam.cx, cam.cy};
      apriltag_pose_t pose;
      const double err = estimate_tag_pose(&info, &pose);
      std::printf("pose t=(%.4f,%.4f,%.4f) err=%.3e\n", pose.t->data[0], pose.t->data[1], pose.t->data[2], err);
      if (!(pose.t->data[2] > 0.1 && pose.t->data[2] < 20.0) || !(err < 1e-3)) rc = 7;  // in front of the camera, metres
      matd_destroy(pose.R);
      matd_destroy(pose.t);
    }
    std::vector<uint8_t> g2(gray.size());
    detector.CopyGrayTo(g2.data());
    if (g2 != gray) rc = 4;
    // a second frame through the same detector, then ReinitializeDetections (apriltags_cuda_detector.cu:497)
    detector.Detect(yuyv.data());
    if (zarray_size(detector.Detections()) != expected) rc = 5;
    detector.ReinitializeDetections();
    if (zarray_size(detector.Detections()) != 0) rc = 6;
  }
  if (argc > 6) {  // the same frame as a JPEG bitstream (camera wire format) through DetectMjpg
    std::vector<uint8_t> jpg;
    FILE *jf = std::fopen(argv[6], "rb");
    if (!jf) return 3;
    uint8_t buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), jf)) > 0) jpg.insert(jpg.end(), buf, buf + n);
    std::fclose(jf);
    frc971::apriltag::GpuDetector detector(width, height, td, cam, dist, B200TAG_FMT_GRAY8);
    detector.DetectMjpg(jpg.data(), jpg.size());
    const zarray_t *detections = detector.Detections();
    std::printf("mjpg detections=%d\n", zarray_size(detections));
    if (zarray_size(detections) != expected) rc = 8;
    for (int i = 0; i < zarray_size(detections); i++) {
      apriltag_detection_t *det;
      zarray_get(detections, i, &det);
      if (det->id != std::atoi(argv[5])) rc = 8;
    }
  }
  apriltag_detector_destroy(td);
  tag36h11_destroy(tf);
  return rc;
}
