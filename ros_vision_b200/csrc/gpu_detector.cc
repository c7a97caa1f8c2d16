// frc971::apriltag::GpuDetector over the C ABI (include/b200tag.h).
// Reference: src/apriltags_cuda/src/apriltag_gpu.cu:111-220 (ctor/dtor/ReinitializeDetections),
// :725-1166 (Detect), src/apriltags_cuda/src/apriltag_detect.cu:94-96,243-258,618-663.
#include "apriltags_cuda/apriltag_gpu.h"

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

// The libapriltag record layouts this layer was written against (upstream AprilTag 3.x, LP64).  In this repo they come
// from the stand-in include/apriltag_compat/apriltag.h; in the real workspace this file is compiled against
// libapriltag's own headers and these assertions make a differing fork fail at compile time
// instead of silently reading the wrong bytes.  -DB200TAG_SKIP_LAYOUT_CHECKS turns them off.
#ifndef B200TAG_SKIP_LAYOUT_CHECKS
#include <cstddef>
static_assert(offsetof(zarray_t, el_sz) == 0 && offsetof(zarray_t, size) == 8 && offsetof(zarray_t, alloc) == 12 &&
                  offsetof(zarray_t, data) == 16 && sizeof(zarray_t) == 24, "zarray_t layout");
static_assert(offsetof(matd_t, nrows) == 0 && offsetof(matd_t, ncols) == 4 && offsetof(matd_t, data) == 8, "matd_t layout");
static_assert(offsetof(apriltag_detection_t, family) == 0 && offsetof(apriltag_detection_t, id) == 8 &&
                  offsetof(apriltag_detection_t, hamming) == 12 && offsetof(apriltag_detection_t, decision_margin) == 16 &&
                  offsetof(apriltag_detection_t, H) == 24 && offsetof(apriltag_detection_t, c) == 32 &&
                  offsetof(apriltag_detection_t, p) == 48 && sizeof(apriltag_detection_t) == 112, "apriltag_detection_t layout");
static_assert(offsetof(apriltag_family_t, ncodes) == 0 && offsetof(apriltag_family_t, codes) == 8 &&
                  offsetof(apriltag_family_t, width_at_border) == 16 && offsetof(apriltag_family_t, total_width) == 20 &&
                  offsetof(apriltag_family_t, reversed_border) == 24 && offsetof(apriltag_family_t, nbits) == 28 &&
                  offsetof(apriltag_family_t, bit_x) == 32 && offsetof(apriltag_family_t, bit_y) == 40 &&
                  offsetof(apriltag_family_t, name) == 56, "apriltag_family_t layout");
static_assert(sizeof(static_cast<apriltag_detector_t *>(nullptr)->quad_decimate) == sizeof(float) &&
                  sizeof(static_cast<apriltag_detector_t *>(nullptr)->decode_sharpening) == sizeof(double) &&
                  sizeof(static_cast<apriltag_detector_t *>(nullptr)->qtp.cos_critical_rad) == sizeof(float),
              "apriltag_detector_t field types");
#endif

namespace frc971::apriltag {
namespace {
bool g_keep_debug_stages = false;

[[noreturn]] void Fatal(const char *what, const char *detail) {
  // the reference aborts through glog LOG(FATAL) / CHECK (cuda_frc971.h:14-17)
  std::fprintf(stderr, "GpuDetector: %s: %s\n", what, detail ? detail : "");
  std::abort();
}

}  // namespace

GpuDetector::GpuDetector(size_t width, size_t height, apriltag_detector_t *tag_detector, CameraMatrix camera_matrix,
                         DistCoeffs distortion_coefficients)
    : width_(width), height_(height), tag_detector_(tag_detector), camera_matrix_(camera_matrix),
      distortion_coefficients_(distortion_coefficients) {
  Init(width, height, tag_detector, camera_matrix, distortion_coefficients, B200TAG_FMT_YUYV);
}

GpuDetector::GpuDetector(size_t width, size_t height, apriltag_detector_t *tag_detector, CameraMatrix camera_matrix,
                         DistCoeffs distortion_coefficients, int pixel_format)
    : width_(width), height_(height), tag_detector_(tag_detector), camera_matrix_(camera_matrix),
      distortion_coefficients_(distortion_coefficients) {
  Init(width, height, tag_detector, camera_matrix, distortion_coefficients, pixel_format);
}

void GpuDetector::Init(size_t width, size_t height, apriltag_detector_t *td, CameraMatrix cam, DistCoeffs dist, int fmt) {
  if (td == nullptr) Fatal("constructor", "tag_detector is null");
  b200tag_config cfg;
  b200tag_default_config(&cfg, static_cast<int>(width), static_cast<int>(height), fmt);
  // apriltag_gpu.cu:166-167: CHECK_EQ(quad_decimate, 2); CHECK(!deglitch).  Integer factors are accepted here.
  const float qd = td->quad_decimate;
  if (qd < 1.0f || qd != static_cast<float>(static_cast<int>(qd))) Fatal("constructor", "quad_decimate must be an integer >= 1");
  if (td->qtp.deglitch) Fatal("constructor", "qtp.deglitch is not supported (apriltag_gpu.cu:167)");
  cfg.quad_decimate = static_cast<int>(qd);
  cfg.quad_sigma = td->quad_sigma;
  cfg.refine_edges = td->refine_edges ? 1 : 0;
  cfg.decode_sharpening = td->decode_sharpening;
  cfg.min_cluster_pixels = td->qtp.min_cluster_pixels;
  cfg.max_nmaxima = td->qtp.max_nmaxima;
  cfg.cos_critical_rad = td->qtp.cos_critical_rad;
  cfg.max_line_fit_mse = td->qtp.max_line_fit_mse;
  cfg.min_white_black_diff = td->qtp.min_white_black_diff;
  if (const char *e = std::getenv("B200TAG_KEEP_STAGES")) g_keep_debug_stages = g_keep_debug_stages || e[0] == '1';
  cfg.keep_stages = g_keep_debug_stages ? 1 : 0;
  cfg.fx = cam.fx; cfg.cx = cam.cx; cfg.fy = cam.fy; cfg.cy = cam.cy;
  cfg.k1 = dist.k1; cfg.k2 = dist.k2; cfg.p1 = dist.p1; cfg.p2 = dist.p2; cfg.k3 = dist.k3;
  // apriltag_gpu.cu:169-177: the families decide border polarity and the minimum tag width; the decoder reads their
  // codes and bit layout.  Every family the caller added is handed over as it is (the eight of apriltag_utils.cu:10-27
  // or any other of at most 64 bits); mixed border polarities abort like the reference's CHECK (apriltag_detect.cu:108).
  if (td->tag_families == nullptr || zarray_size(td->tag_families) < 1) Fatal("constructor", "the detector has no tag family");
  std::vector<b200tag_family> fams;
  for (int i = 0; i < zarray_size(td->tag_families); i++) {
    apriltag_family_t *family;
    zarray_get(td->tag_families, i, &family);
    b200tag_family f;
    f.name = family->name;
    f.nbits = family->nbits;
    f.ncodes = family->ncodes;
    f.codes = family->codes;
    f.bit_x = family->bit_x;
    f.bit_y = family->bit_y;
    f.width_at_border = family->width_at_border;
    f.total_width = family->total_width;
    f.reversed_border = family->reversed_border ? 1 : 0;
    f.max_hamming = 2;  // apriltag_detector_add_family (bits_corrected = 2), apriltags_cuda_detector.cu:140
    fams.push_back(f);
  }
  const int rc = b200tag_create_families(&cfg, fams.data(), static_cast<int>(fams.size()), &handle_);
  if (rc != 0) Fatal(b200tag_error_string(rc), b200tag_last_error(nullptr));
  detections_ = zarray_create(sizeof(apriltag_detection_t *));
  zarray_ensure_capacity(detections_, static_cast<int>(kMaxBlobs));
}

void GpuDetector::KeepDebugStages(bool keep) { g_keep_debug_stages = keep; }

GpuDetector::~GpuDetector() {
  ClearDetections();
  zarray_destroy(detections_);
  b200tag_destroy(handle_);
}

void GpuDetector::ClearDetections() {
  for (int i = 0; i < zarray_size(detections_); ++i) {
    apriltag_detection_t *det;
    zarray_get(detections_, i, &det);
    apriltag_detection_destroy(det);
  }
  zarray_truncate(detections_, 0);
}

void GpuDetector::ReinitializeDetections() {  // apriltag_gpu.cu:202-220
  ClearDetections();
  zarray_destroy(detections_);
  detections_ = zarray_create(sizeof(apriltag_detection_t *));
  zarray_ensure_capacity(detections_, static_cast<int>(kMaxBlobs));
}

void GpuDetector::Detect(const uint8_t *image) { Collect(b200tag_detect(handle_, image)); }

void GpuDetector::DetectMjpg(const uint8_t *jpeg, size_t size) {
  const uint8_t *jpegs[1] = {jpeg};
  const size_t sizes[1] = {size};
  Collect(b200tag_detect_mjpg(handle_, jpegs, sizes, 1));
}

void GpuDetector::Collect(int rc) {
  if (rc != 0 && rc != B200TAG_E_OVERFLOW) Fatal(b200tag_error_string(rc), b200tag_last_error(handle_));
  if (rc == B200TAG_E_OVERFLOW) std::fprintf(stderr, "GpuDetector: %s\n", b200tag_last_error(handle_));
  // DecodeTags (apriltag_detect.cu:626-632): previous detections are destroyed first
  ClearDetections();
  quad_corners_host_.clear();
  int n = 0;
  const b200tag_detection *d = b200tag_detections(handle_, 0, &n);
  for (int i = 0; i < n; i++) {  // already reconciled and sorted by id (apriltag_detect.cu:660-662)
    apriltag_detection_t *det = static_cast<apriltag_detection_t *>(calloc(1, sizeof(apriltag_detection_t)));
    apriltag_family_t *family = nullptr;
    zarray_get(tag_detector_->tag_families, d[i].family, &family);
    det->family = family;
    det->id = d[i].id;
    det->hamming = d[i].hamming;
    det->decision_margin = d[i].decision_margin;
    det->H = matd_create_data(3, 3, d[i].H);
    det->c[0] = d[i].c[0];
    det->c[1] = d[i].c[1];
    std::memcpy(det->p, d[i].p, sizeof(det->p));
    zarray_add(detections_, &det);
  }
}

const std::vector<QuadCorners> &GpuDetector::FitQuads() const {  // apriltag_detect.cu:94-96
  int n = 0;
  const b200tag_quad *q = b200tag_quads(handle_, 0, &n);
  quad_corners_host_.resize(static_cast<size_t>(n));
  for (int i = 0; i < n; i++) {
    std::memcpy(quad_corners_host_[i].corners, q[i].corners, sizeof(q[i].corners));
    quad_corners_host_[i].reversed_border = q[i].reversed_border != 0;
    quad_corners_host_[i].blob_index = q[i].blob_index;
  }
  return quad_corners_host_;
}

namespace {
void CopyStage(b200tag_detector *h, int stage, void *out) {
  size_t bytes = 0;
  if (b200tag_copy_stage(h, 0, stage, nullptr, 0, &bytes) != 0) Fatal("copy_stage", b200tag_last_error(h));
  if (b200tag_copy_stage(h, 0, stage, out, bytes, &bytes) != 0) Fatal("copy_stage", b200tag_last_error(h));
}
}  // namespace

void GpuDetector::CopyGrayTo(uint8_t *output) const { CopyStage(handle_, B200TAG_STAGE_GRAY, output); }
void GpuDetector::CopyDecimatedTo(uint8_t *output) const { CopyStage(handle_, B200TAG_STAGE_QUAD_IMAGE, output); }
void GpuDetector::CopyThresholdedTo(uint8_t *output) const { CopyStage(handle_, B200TAG_STAGE_THRESHOLD, output); }
void GpuDetector::CopyUnionMarkersTo(uint32_t *output) const { CopyStage(handle_, B200TAG_STAGE_LABELS, output); }
void GpuDetector::CopyUnionMarkersSizeTo(uint32_t *output) const { CopyStage(handle_, B200TAG_STAGE_SIZES, output); }

int GpuDetector::NumCompressedUnionMarkerPairs() const {
  b200tag_frame_info info;
  return b200tag_frame_info_get(handle_, 0, &info) == 0 ? static_cast<int>(info.num_points) : 0;
}
int GpuDetector::NumQuads() const {
  b200tag_frame_info info;
  return b200tag_frame_info_get(handle_, 0, &info) == 0 ? static_cast<int>(info.num_clusters) : 0;
}
int GpuDetector::NumSelectedPairs() const {
  b200tag_frame_info info;
  return b200tag_frame_info_get(handle_, 0, &info) == 0 ? static_cast<int>(info.num_selected_points) : 0;
}
int GpuDetector::NumFitQuads() const {
  b200tag_frame_info info;
  return b200tag_frame_info_get(handle_, 0, &info) == 0 ? static_cast<int>(info.num_fit_quads) : 0;
}

void GpuDetector::AdjustCenter(float corners[4][2]) const {  // apriltag_detect.cu:243-258
  const float quad_decimate = tag_detector_->quad_decimate;
  if (tag_detector_->quad_decimate > 1) {
    if (tag_detector_->quad_decimate == 1.5) {
      for (int j = 0; j < 4; j++) {
        corners[j][0] *= quad_decimate;
        corners[j][1] *= quad_decimate;
      }
    } else {
      for (int j = 0; j < 4; j++) {
        corners[j][0] = (corners[j][0] - 0.5f) * quad_decimate + 0.5f;
        corners[j][1] = (corners[j][1] - 0.5f) * quad_decimate + 0.5f;
      }
    }
  }
}

void GpuDetector::SetCameraMatrix(CameraMatrix m) {
  camera_matrix_ = m;
  b200tag_set_camera(handle_, m.fx, m.cx, m.fy, m.cy);
}

void GpuDetector::SetDistortionCoefficients(DistCoeffs d) {
  distortion_coefficients_ = d;
  b200tag_set_distortion(handle_, d.k1, d.k2, d.p1, d.p2, d.k3);
}

bool GpuDetector::UnDistort(double *u, double *v, const CameraMatrix *m, const DistCoeffs *d) {
  return b200tag_undistort(u, v, m->fx, m->cx, m->fy, m->cy, d->k1, d->k2, d->p1, d->p2, d->k3) == 1;
}

}  // namespace frc971::apriltag

// libapriltag's estimate_tag_pose, for builds that link this library instead of libapriltag
// (apriltags_cuda_detector.cu:425-462).  With the real libapriltag on the link line its own definition wins.
extern "C" __attribute__((weak)) double estimate_tag_pose(apriltag_detection_info_t *info, apriltag_pose_t *pose) {
  b200tag_detection d;
  std::memset(&d, 0, sizeof(d));
  const apriltag_detection_t *det = info->det;
  d.id = det->id;
  for (int i = 0; i < 9; i++) d.H[i] = det->H->data[i];
  for (int i = 0; i < 4; i++) {
    d.p[i][0] = det->p[i][0];
    d.p[i][1] = det->p[i][1];
  }
  d.c[0] = det->c[0];
  d.c[1] = det->c[1];
  b200tag_pose out;
  if (b200tag_estimate_pose(&d, info->tagsize, info->fx, info->fy, info->cx, info->cy, &out) != 0) return HUGE_VAL;
  pose->R = matd_create_data(3, 3, out.R);
  pose->t = matd_create_data(3, 1, out.t);
  return out.err;
}
