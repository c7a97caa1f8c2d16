// Exercises the typed debug accessors of frc971::apriltag::GpuDetector (reference: apriltag_gpu.h:111-183) on one gray
// frame and prints order-independent summaries that tests/test_cpp_class.py compares with the CPU oracle:
//   debug_accessors_test <gray.raw> <width> <height>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "apriltags_cuda/apriltag_gpu.h"

using namespace frc971::apriltag;

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int width = std::atoi(argv[2]), height = std::atoi(argv[3]);
  std::vector<uint8_t> gray(static_cast<size_t>(width) * height);
  FILE *f = std::fopen(argv[1], "rb");
  if (!f || std::fread(gray.data(), 1, gray.size(), f) != gray.size()) return 3;
  std::fclose(f);
  apriltag_family_t *tf = tag36h11_create();
  apriltag_detector_t *td = apriltag_detector_create();
  apriltag_detector_add_family(td, tf);
  td->quad_decimate = 2.0;
  int rc = 0;
  {
    GpuDetector::KeepDebugStages(true);
    GpuDetector det(width, height, td, CameraMatrix{1, 0, 1, 0}, DistCoeffs{0, 0, 0, 0, 0}, B200TAG_FMT_GRAY8);
    det.Detect(gray.data());
    const int w = width / 2, h = height / 2;
    const int np = det.NumCompressedUnionMarkerPairs();
    std::vector<QuadBoundaryPoint> dense(static_cast<size_t>(4) * (w - 2) * (h - 2)), comp(np);
    det.CopyUnionMarkerPairTo(dense.data());
    det.CopyCompressedUnionMarkerPairTo(comp.data());
    size_t nonzero = 0, ci = 0;
    bool dense_matches = true;
    for (const QuadBoundaryPoint &q : dense)
      if (q.nonzero()) {
        nonzero++;
        dense_matches = dense_matches && ci < comp.size() && comp[ci] == q;
        ci++;
      }
    const std::vector<QuadBoundaryPoint> sorted = det.CopySortedUnionMarkerPair();
    bool monotone = true;
    uint64_t xsum = 0, ysum = 0;
    for (size_t i = 0; i < sorted.size(); i++) {
      if (i && sorted[i - 1].rep01() > sorted[i].rep01()) monotone = false;
      xsum += sorted[i].x();
      ysum += sorted[i].y();
    }
    std::printf("points=%d dense_nonzero=%zu dense_matches=%d sorted=%zu monotone=%d xsum=%" PRIu64 " ysum=%" PRIu64 "\n", np, nonzero,
                dense_matches ? 1 : 0, sorted.size(), monotone ? 1 : 0, xsum, ysum);
    const std::vector<MinMaxExtents> ext = det.CopyExtents();
    uint64_t count_sum = 0;
    int64_t dot_sum = 0;
    bool offsets_ok = true;
    for (size_t i = 0; i < ext.size(); i++) {
      if (ext[i].starting_offset != count_sum) offsets_ok = false;
      count_sum += ext[i].count;
      dot_sum += ext[i].pxgx_plus_pygy_sum;
      // every point of the pair lies inside its box
      for (uint32_t k = 0; k < ext[i].count; k++) {
        const QuadBoundaryPoint &q = sorted[ext[i].starting_offset + k];
        if (q.x() < ext[i].min_x || q.x() > ext[i].max_x || q.y() < ext[i].min_y || q.y() > ext[i].max_y) offsets_ok = false;
      }
    }
    std::printf("pairs=%zu numquads=%d count_sum=%" PRIu64 " dot_sum=%" PRId64 " offsets_ok=%d\n", ext.size(), det.NumQuads(), count_sum,
                dot_sum, offsets_ok ? 1 : 0);
    const auto sel = det.CopySelectedExtents();
    size_t nsel = 0, sel_points = 0;
    for (const auto &kv : sel)
      if (kv.value.count) {
        nsel++;
        sel_points += kv.value.count;
      }
    const std::vector<IndexPoint> unsorted = det.CopySelectedBlobs(), bysort = det.CopySortedSelectedBlobs();
    bool theta_ok = bysort.size() == unsorted.size();
    uint64_t theta_sum = 0, theta_sum2 = 0;
    for (size_t i = 0; i < bysort.size(); i++) {
      if (i && bysort[i - 1].blob_index() == bysort[i].blob_index() && bysort[i - 1].theta() > bysort[i].theta()) theta_ok = false;
      if (i && bysort[i - 1].blob_index() > bysort[i].blob_index()) theta_ok = false;
      theta_sum += bysort[i].theta();
    }
    for (const IndexPoint &p : unsorted) theta_sum2 += p.theta();
    std::printf("selected_pairs=%zu selected_points=%zu numselected=%d index_points=%zu theta_ok=%d theta_sum=%" PRIu64 " same_sum=%d\n", nsel,
                sel_points, det.NumSelectedPairs(), bysort.size(), theta_ok ? 1 : 0, theta_sum, theta_sum == theta_sum2 ? 1 : 0);
    const std::vector<LineFitPoint> lfp = det.CopyLineFitPoints();
    const std::vector<double> errs = det.CopyErrors(), filt = det.CopyFilteredErrors();
    int64_t w_last = 0;
    for (size_t i = 0; i < lfp.size(); i++)
      if (i + 1 == lfp.size() || lfp[i + 1].blob_index != lfp[i].blob_index) w_last += lfp[i].W;
    double esum = 0, fsum = 0;
    for (double e : errs) esum += e;
    for (double e : filt) fsum += e;
    std::printf("lfp=%zu w_last=%" PRId64 " errs=%zu esum=%.6f filt=%zu fsum=%.6f\n", lfp.size(), w_last, errs.size(), esum, filt.size(), fsum);
    const std::vector<Peak> peaks = det.CopyPeaks(), cpeaks = det.CopyCompressedPeaks();
    size_t npk = 0;
    for (const Peak &p : peaks) npk += p.blob_index != Peak::kNoPeak();
    bool peaks_sorted = true;
    for (size_t i = 1; i < cpeaks.size(); i++)
      if (cpeaks[i - 1].blob_index > cpeaks[i].blob_index ||
          (cpeaks[i - 1].blob_index == cpeaks[i].blob_index && cpeaks[i - 1].error > cpeaks[i].error))
        peaks_sorted = false;
    const std::vector<FitQuad> fq = det.CopyFitQuads();
    size_t nvalid = 0;
    for (const FitQuad &q : fq) nvalid += q.valid;
    std::printf("peaks=%zu is_peak=%zu compressed=%d sorted=%d fitquads=%zu numfitquads=%d valid=%zu detections=%d\n", peaks.size(), npk,
                det.NumCompressedPeaks(), peaks_sorted ? 1 : 0, fq.size(), det.NumFitQuads(), nvalid, zarray_size(det.Detections()));
    if (!dense_matches || !monotone || !offsets_ok || !theta_ok || !peaks_sorted) rc = 1;
  }
  apriltag_detector_destroy(td);
  tag36h11_destroy(tf);
  return rc;
}
