// TEST INFRASTRUCTURE -- C-ABI harness around the REFERENCE's own GpuDetector.
//
// oracle/build_ref.sh compiles the reference's CUDA sources where they lie under
// /root/reference/src/apriltags_cuda (never copied into this repository) together with this file
// into oracle/_ref/librefgpu.so.  tests/test_gpu_reference_live.py loads it on the B200 box, runs
// frc971::apriltag::GpuDetector::Detect (apriltag_gpu.cu:725-1166) on the same YUYV frames as the
// product and the CPU oracle, and compares every stage the reference exposes through its debug
// accessors (apriltag_gpu.h:97-183).
//
// libapriltag is not vendored in the reference tree, so its entry points are defined here:
//   * workerpool_*: tasks run sequentially in workerpool_run;
//   * quad_decode_index (declared at apriltag_detect.cu:27-29): records the quad AFTER the
//     reference's own RefineEdges (apriltag_detect.cu:405-564) ran on it, produces no detection;
//   * reconcile_detections: no-op.
// The harness therefore pins the path through the refined quad corners; decode stays
// "parity unpinned" (DESIGN.md section 5).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "apriltags_cuda/apriltag_gpu.h"

namespace {
struct Task {
  void (*f)(void *);
  void *p;
};
std::vector<Task> g_tasks;
struct RefinedQuad {
  float p[4][2];
  int32_t reversed_border;
};
std::vector<RefinedQuad> g_refined;
}  // namespace

extern "C" {

void workerpool_add_task(workerpool_t *, void (*f)(void *p), void *p) { g_tasks.push_back(Task{f, p}); }
void workerpool_run(workerpool_t *) {
  for (const Task &t : g_tasks) t.f(t.p);
  g_tasks.clear();
}
image_u8_t *image_u8_copy(const image_u8_t *) { return nullptr; }
void image_u8_darken(image_u8_t *) {}
void image_u8_draw_line(image_u8_t *, float, float, float, float, int, int) {}
int image_u8_write_pnm(const image_u8_t *, const char *) { return 0; }

zarray_t *g2d_polygon_create_zeros(int sz) {
  zarray_t *points = zarray_create(sizeof(double[2]));
  double z[2] = {0, 0};
  for (int i = 0; i < sz; i++) zarray_add(points, &z);
  return points;
}

void quad_decode_index(apriltag_detector_t *, struct quad *q, image_u8_t *, image_u8_t *, zarray_t *) {
  RefinedQuad r;
  std::memcpy(r.p, q->p, sizeof(r.p));
  r.reversed_border = q->reversed_border;
  g_refined.push_back(r);
}

void reconcile_detections(zarray_t *, zarray_t *, zarray_t *) {}

// ---------------------------------------------------------------------------------------------
// flat C API
// ---------------------------------------------------------------------------------------------
struct refgpu {
  apriltag_detector_t *td;
  apriltag_family_t *tf;
  frc971::apriltag::GpuDetector *det;
  int width, height;
};

// Unpacked records; layouts mirror oracle/apriltag_oracle.h so the test compares arrays directly.
struct refgpu_point {
  uint32_t rep0, rep1;
  uint16_t x, y, bx, by;
  uint8_t dir, b2w, pad[2];
};
struct refgpu_extents {
  uint16_t min_x, min_y, max_x, max_y;
  uint32_t start, count;
  int32_t gx_sum, gy_sum;
  int64_t pxgx_plus_pygy_sum;
};
struct refgpu_spoint {
  uint32_t blob, theta;
  uint16_t x, y, bx, by;
  uint8_t dir, b2w, pad[2];
};
struct refgpu_lfp {
  int64_t Mxx, Myy, Mxy, Mx, My, W;
  uint32_t blob, pad;
};
struct refgpu_moments {
  int64_t Mx, My, W, Mxx, Myy, Mxy;
  int32_t N, pad;
};
struct refgpu_fitquad {
  uint32_t blob;
  int32_t valid;
  uint32_t indices[4];
  refgpu_moments moments[4];
};
struct refgpu_corners {
  float corners[4][2];
  int32_t reversed_border;
  uint32_t blob;
};

refgpu *refgpu_create(int width, int height, const double cam[4], const double dist[5], int min_cluster_pixels,
                      int refine_edges) {
  refgpu *r = new refgpu();
  r->width = width;
  r->height = height;
  r->td = apriltag_detector_create();
  r->tf = tag36h11_create();
  apriltag_detector_add_family_bits(r->td, r->tf, 2);
  r->td->quad_decimate = 2.0f;
  r->td->quad_sigma = 0.0f;
  r->td->nthreads = 1;
  r->td->debug = false;
  r->td->refine_edges = refine_edges != 0;
  r->td->qtp.min_cluster_pixels = min_cluster_pixels;
  if (!r->td->wp) r->td->wp = workerpool_create(1);
  frc971::apriltag::CameraMatrix cm{cam[0], cam[1], cam[2], cam[3]};
  frc971::apriltag::DistCoeffs dc{dist[0], dist[1], dist[2], dist[3], dist[4]};
  r->det = new frc971::apriltag::GpuDetector(width, height, r->td, cm, dc);
  return r;
}

void refgpu_destroy(refgpu *r) {
  delete r->det;
  apriltag_detector_destroy(r->td);
  tag36h11_destroy(r->tf);
  delete r;
}

void refgpu_detect(refgpu *r, const uint8_t *yuyv) {
  g_refined.clear();
  r->det->Detect(yuyv);
}

// Wall-clock milliseconds per synchronous Detect call (H2D copy from the caller's pageable buffer
// included, exactly as the node calls it), decode stubbed.
double refgpu_time_detect(refgpu *r, const uint8_t *const *frames, int nframes, int iters, int warmup) {
  for (int i = 0; i < warmup; i++) refgpu_detect(r, frames[i % nframes]);
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) refgpu_detect(r, frames[i % nframes]);
  cudaDeviceSynchronize();
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
}

void refgpu_copy_gray(refgpu *r, uint8_t *out) { r->det->CopyGrayTo(out); }
void refgpu_copy_decimated(refgpu *r, uint8_t *out) { r->det->CopyDecimatedTo(out); }
void refgpu_copy_thresholded(refgpu *r, uint8_t *out) { r->det->CopyThresholdedTo(out); }
void refgpu_copy_labels(refgpu *r, uint32_t *out) { r->det->CopyUnionMarkersTo(out); }
void refgpu_copy_sizes(refgpu *r, uint32_t *out) { r->det->CopyUnionMarkersSizeTo(out); }

int refgpu_num_points(refgpu *r) { return r->det->NumCompressedUnionMarkerPairs(); }
// Boundary points in the reference's order after its blob-pair sort (apriltag_gpu.cu:813-825).
int refgpu_copy_sorted_points(refgpu *r, refgpu_point *out, int cap) {
  const auto v = r->det->CopySortedUnionMarkerPair();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++) {
    const auto &q = v[i];
    out[i] = refgpu_point{q.rep0(), q.rep1(), static_cast<uint16_t>(q.x()), static_cast<uint16_t>(q.y()),
                          static_cast<uint16_t>(q.base_x()), static_cast<uint16_t>(q.base_y()),
                          static_cast<uint8_t>(q.key & 3), static_cast<uint8_t>(q.black_to_white()), {0, 0}};
  }
  return static_cast<int>(v.size());
}

int refgpu_num_blob_pairs(refgpu *r) { return r->det->NumQuads(); }
int refgpu_copy_extents(refgpu *r, refgpu_extents *out, int cap) {
  const auto v = r->det->CopyExtents();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++)
    out[i] = refgpu_extents{v[i].min_x, v[i].min_y, v[i].max_x, v[i].max_y, v[i].starting_offset, v[i].count,
                            v[i].gx_sum, v[i].gy_sum, v[i].pxgx_plus_pygy_sum};
  return static_cast<int>(v.size());
}
// Extents after SelectBlobs zeroed the rejected ones and rebased the offsets (apriltag_gpu.cu:873-905).
int refgpu_copy_selected_extents(refgpu *r, refgpu_extents *out, int cap) {
  const auto v = r->det->CopySelectedExtents();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++) {
    const auto &e = v[i].value;
    out[i] = refgpu_extents{e.min_x, e.min_y, e.max_x, e.max_y, e.starting_offset, e.count, e.gx_sum, e.gy_sum,
                            e.pxgx_plus_pygy_sum};
  }
  return static_cast<int>(v.size());
}

int refgpu_num_selected_points(refgpu *r) { return r->det->NumSelectedPairs(); }
// Selected points after the (blob, theta) sort (apriltag_gpu.cu:944-956).
int refgpu_copy_sorted_selected(refgpu *r, refgpu_spoint *out, int cap) {
  const auto v = r->det->CopySortedSelectedBlobs();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++) {
    const auto &q = v[i];
    out[i] = refgpu_spoint{q.blob_index(), q.theta(), static_cast<uint16_t>(q.x()), static_cast<uint16_t>(q.y()),
                           static_cast<uint16_t>(q.base_x()), static_cast<uint16_t>(q.base_y()),
                           static_cast<uint8_t>(q.key & 3), static_cast<uint8_t>(q.black_to_white()), {0, 0}};
  }
  return static_cast<int>(v.size());
}
int refgpu_copy_line_fit_points(refgpu *r, refgpu_lfp *out, int cap) {
  const auto v = r->det->CopyLineFitPoints();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++)
    out[i] = refgpu_lfp{v[i].Mxx, v[i].Myy, v[i].Mxy, v[i].Mx, v[i].My, v[i].W, v[i].blob_index, 0};
  return static_cast<int>(v.size());
}
int refgpu_copy_errors(refgpu *r, double *errs, double *filtered, int cap) {
  const auto e = r->det->CopyErrors();
  const auto f = r->det->CopyFilteredErrors();
  const int n = static_cast<int>(e.size()) < cap ? static_cast<int>(e.size()) : cap;
  std::memcpy(errs, e.data(), n * sizeof(double));
  std::memcpy(filtered, f.data(), n * sizeof(double));
  return static_cast<int>(e.size());
}
int refgpu_num_fit_quads(refgpu *r) { return r->det->NumFitQuads(); }
int refgpu_copy_fit_quads(refgpu *r, refgpu_fitquad *out, int cap) {
  const auto v = r->det->CopyFitQuads();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++) {
    out[i].blob = v[i].blob_index;
    out[i].valid = v[i].valid;
    for (int k = 0; k < 4; k++) {
      out[i].indices[k] = v[i].indices[k];
      const auto &m = v[i].moments[k];
      out[i].moments[k] = refgpu_moments{m.Mx, m.My, m.W, m.Mxx, m.Myy, m.Mxy, m.N, 0};
    }
  }
  return static_cast<int>(v.size());
}
// QuadCorners after UpdateFitQuads + AdjustPixelCenters (apriltag_detect.cu:38-282).
int refgpu_copy_quad_corners(refgpu *r, refgpu_corners *out, int cap) {
  const auto &v = r->det->FitQuads();
  const int n = static_cast<int>(v.size()) < cap ? static_cast<int>(v.size()) : cap;
  for (int i = 0; i < n; i++) {
    std::memcpy(out[i].corners, v[i].corners, sizeof(out[i].corners));
    out[i].reversed_border = v[i].reversed_border;
    out[i].blob = v[i].blob_index;
  }
  return static_cast<int>(v.size());
}
// Quads as handed to quad_decode_index, i.e. after the reference's RefineEdges; same order as
// refgpu_copy_quad_corners.
int refgpu_copy_refined_quads(refgpu *, refgpu_corners *out, int cap) {
  const int n = static_cast<int>(g_refined.size()) < cap ? static_cast<int>(g_refined.size()) : cap;
  for (int i = 0; i < n; i++) {
    std::memcpy(out[i].corners, g_refined[i].p, sizeof(out[i].corners));
    out[i].reversed_border = g_refined[i].reversed_border;
    out[i].blob = static_cast<uint32_t>(i);
  }
  return static_cast<int>(g_refined.size());
}

}  // extern "C"
