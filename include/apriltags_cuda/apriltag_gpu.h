// Header-compatible replacement for the reference's
// src/apriltags_cuda/include/apriltags_cuda/apriltag_gpu.h (frc971::apriltag::GpuDetector, :77-359).
//
// Same namespace, class name, public members and ownership rules, so
// ApriltagsDetector::setup_apriltags / imageCallback (apriltags_cuda_detector.cu:178-179,404,410,497),
// gpu_detector_test.cu and the demos compile against it unchanged.  The implementation forwards to the
// C ABI in include/b200tag.h (libb200tag.so); nothing CUDA-specific leaks into this header, so callers no
// longer need nvcc/clang-cuda to include it.
//
// The debug copies typed to the reference's packed device structs (QuadBoundaryPoint / IndexPoint / LineFitPoint /
// Peak / FitQuad, apriltag_gpu.h:111-183) are here too: reference_types.h holds the record types, the accessors convert
// the engine's own (wider) stage records on the host.  They need a detector that keeps its debug stages
// (GpuDetector::KeepDebugStages(true) before construction, or B200TAG_KEEP_STAGES=1) and a frame that fits the
// reference's bit fields (quad image <= 1024 x 1024, <= 4096 blob pairs); the same stages are available without those
// limits, unpacked, through b200tag_copy_stage().
#ifndef B200TAG_APRILTAGS_CUDA_APRILTAG_GPU_H_
#define B200TAG_APRILTAGS_CUDA_APRILTAG_GPU_H_

#include <cstddef>
#include <cstdint>
#include <vector>

extern "C" {
#include "apriltag.h"
}
#include "b200tag.h"
#include "apriltags_cuda/reference_types.h"

namespace frc971::apriltag {

struct QuadCorners {  // apriltag_gpu.h:55-59
  float corners[4][2];
  bool reversed_border;
  uint32_t blob_index;
};

struct CameraMatrix {  // apriltag_gpu.h:61-66 (note the field order)
  double fx;
  double cx;
  double fy;
  double cy;
};

struct DistCoeffs {  // apriltag_gpu.h:68-74
  double k1;
  double k2;
  double p1;
  double p2;
  double k3;
};

class GpuDetector {
 public:
  // The number of blobs we will consider when counting april tags (apriltag_gpu.h:80).  The reference
  // silently overflows beyond this; this engine sizes its lists from the frame and reports overflow.
  static constexpr size_t kMaxBlobs = 2048;

  // apriltag_gpu.h:84-85.  `tag_detector` stays owned by the caller and is read at construction
  // (quad_decimate, quad_sigma, refine_edges, decode_sharpening, qtp, tag_families).
  // Violated preconditions abort, like the reference's CHECK/LOG(FATAL) (cuda_frc971.h:14-17).
  GpuDetector(size_t width, size_t height, apriltag_detector_t *tag_detector, CameraMatrix camera_matrix,
              DistCoeffs distortion_coefficients);
  // Extension: frames in another pixel format (B200TAG_FMT_*), e.g. the node's bgr8 image directly.
  GpuDetector(size_t width, size_t height, apriltag_detector_t *tag_detector, CameraMatrix camera_matrix,
              DistCoeffs distortion_coefficients, int pixel_format);
  virtual ~GpuDetector();
  GpuDetector(const GpuDetector &) = delete;
  GpuDetector &operator=(const GpuDetector &) = delete;

  // Detects april tags in the provided image (host pointer, YUYV 4:2:2 unless constructed otherwise).
  void Detect(const uint8_t *image);
  // Extension (camera wire format): the same for one JPEG bitstream of a camera frame -- what the cameras send before
  // the reference's camera node decodes it with OpenCV (camera_publisher.cpp:198,336).  The luminance plane is decoded
  // on the GPU; the detector must have been constructed with pixel format B200TAG_FMT_GRAY8.
  void DetectMjpg(const uint8_t *jpeg, size_t size);

  const std::vector<QuadCorners> &FitQuads() const;
  const zarray_t *Detections() const { return detections_; }
  void ReinitializeDetections();

  // Debug methods to expose internal state for testing (apriltag_gpu.h:98-109,131-133).
  void CopyGrayTo(uint8_t *output) const;
  void CopyDecimatedTo(uint8_t *output) const;
  void CopyThresholdedTo(uint8_t *output) const;
  void CopyUnionMarkersTo(uint32_t *output) const;
  void CopyUnionMarkersSizeTo(uint32_t *output) const;
  int NumCompressedUnionMarkerPairs() const;  // :127
  int NumQuads() const;                       // :135  number of blob pairs
  int NumSelectedPairs() const;               // :146  points of the selected blobs
  int NumFitQuads() const;                    // :179

  // The typed debug copies of apriltag_gpu.h:111-183 (see reference_types.h for their limits).  Blob indices are
  // positions in the list of blob pairs sorted by (rep1, rep0), the order the reference's radix sort produces.
  void CopyUnionMarkerPairTo(QuadBoundaryPoint *output) const;            // :111  dense, 4 (w-2) (h-2) entries
  void CopyCompressedUnionMarkerPairTo(QuadBoundaryPoint *output) const;  // :115  NumCompressedUnionMarkerPairs()
  std::vector<QuadBoundaryPoint> CopySortedUnionMarkerPair() const;       // :119
  std::vector<MinMaxExtents> CopyExtents() const;                         // :137  NumQuads() blob pairs
  std::vector<cub::KeyValuePair<long, MinMaxExtents>> CopySelectedExtents() const;  // :141
  std::vector<IndexPoint> CopySelectedBlobs() const;                      // :148  NumSelectedPairs() points
  std::vector<IndexPoint> CopySortedSelectedBlobs() const;                // :152
  std::vector<LineFitPoint> CopyLineFitPoints() const;                    // :156
  std::vector<double> CopyErrors() const;                                 // :160
  std::vector<double> CopyFilteredErrors() const;                         // :164
  std::vector<Peak> CopyPeaks() const;                                    // :167
  int NumCompressedPeaks() const;                                         // :171
  std::vector<Peak> CopyCompressedPeaks() const;                          // :175
  std::vector<FitQuad> CopyFitQuads() const;                              // :181
  // Detectors constructed after KeepDebugStages(true) keep every intermediate stage (slower: extra kernels and
  // copies); the typed accessors above abort on a detector that does not.
  static void KeepDebugStages(bool keep);

  void AdjustCenter(float corners[4][2]) const;  // :185

  void SetCameraMatrix(CameraMatrix camera_matrix);                       // :189-191
  void SetDistortionCoefficients(DistCoeffs distortion_coefficients);     // :193-195

  // Undistort pixels based on our camera model, using iterative algorithm.  Returns false if we fail
  // to converge (:199-200).
  static bool UnDistort(double *u, double *v, const CameraMatrix *camera_matrix,
                        const DistCoeffs *distortion_coefficients);

  // The underlying C-ABI handle (batching, stage copies, profiling).
  b200tag_detector *handle() const { return handle_; }

 private:
  void Init(size_t width, size_t height, apriltag_detector_t *td, CameraMatrix cam, DistCoeffs dist, int fmt);
  void ClearDetections();
  void Collect(int rc);  // result of a b200tag_detect* call -> detections_ (DecodeTags, apriltag_detect.cu:626-662)

  const size_t width_;
  const size_t height_;
  apriltag_detector_t *tag_detector_;
  b200tag_detector *handle_ = nullptr;
  CameraMatrix camera_matrix_;
  DistCoeffs distortion_coefficients_;
  mutable std::vector<QuadCorners> quad_corners_host_;
  zarray_t *detections_ = nullptr;
};

}  // namespace frc971::apriltag

#endif
