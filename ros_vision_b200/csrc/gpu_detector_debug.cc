// The typed debug copies of frc971::apriltag::GpuDetector (reference: apriltag_gpu.h:111-183): the engine's stage
// records (b200tag_copy_stage) converted on the host into the reference's packed records (reference_types.h), in the
// orders the reference's pipeline produces them:
//   dense boundary array      index = (w-2)(h-2) * dir + (x-1) + (y-1)(w-2)                (apriltag_gpu.cu:276-322)
//   compressed                the non-zero entries of the dense array in index order     (:788-802)
//   sorted                    stable by (rep1, rep0)                                       (:813-825)
//   extents                   one per blob pair in that order                              (:829-862)
//   selected blobs / points   blob index = position of the pair in the extents list       (:380-412,873-956)
//   peaks                     per point; compressed = local maxima by (blob, -error)       (:1001-1078)
// Host-side and debug-only: nothing here is on the detection path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "apriltags_cuda/apriltag_gpu.h"

namespace frc971::apriltag {
namespace {

[[noreturn]] void Die(const char *what) {
  std::fprintf(stderr, "GpuDetector (typed debug accessor): %s\n", what);
  std::abort();
}

template <typename T>
std::vector<T> Stage(b200tag_detector *h, int stage) {
  size_t bytes = 0;
  if (b200tag_copy_stage(h, 0, stage, nullptr, 0, &bytes) != 0)
    Die("this stage is not kept: construct the detector after GpuDetector::KeepDebugStages(true) (or B200TAG_KEEP_STAGES=1)");
  std::vector<T> out(bytes / sizeof(T));
  if (bytes && b200tag_copy_stage(h, 0, stage, out.data(), bytes, &bytes) != 0) Die(b200tag_last_error(h));
  return out;
}

struct Snapshot {
  int w = 0, h = 0;                      // quad image
  std::vector<b200tag_blob> pairs;       // every blob pair, sorted by (rep1, rep0)
  std::vector<uint32_t> first_point;     // per pair: first position in the sorted point list
  std::vector<QuadBoundaryPoint> sorted; // all boundary points, sorted (stable) by pair
  std::vector<QuadBoundaryPoint> compressed;
  // selected pairs only, in pair order
  struct Sel {
    uint32_t pair;      // index into pairs
    uint32_t seg;       // first record of the blob in the engine's per-blob arrays
    uint32_t count;
    uint32_t first;     // first position among the selected points
  };
  std::vector<Sel> sel;
  std::vector<uint64_t> keys;
  std::vector<b200tag_lfp> lfp;
  std::vector<float> errs;
  std::vector<double> filt;
  std::vector<b200tag_blob> blobs;       // candidate blobs as the engine numbers them (FitQuad::blob_index)
  std::vector<uint32_t> pair_of_blob;    // engine blob -> pair index
};

uint64_t PairKey(uint32_t rep0, uint32_t rep1) { return (static_cast<uint64_t>(rep1) << 32) | rep0; }

Snapshot Take(const GpuDetector &det, int w, int h, bool points, bool per_blob) {
  Snapshot s;
  s.w = w;
  s.h = h;
  if (w > 1024 || h > 1024) Die("quad image larger than 1024 x 1024: does not fit the reference's 10-bit coordinates (points.h:62-74)");
  if (static_cast<size_t>(w) * h > (1u << 20)) Die("labels do not fit the reference's 20-bit blob ids (points.h:32-49)");
  b200tag_detector *hd = det.handle();
  s.pairs = Stage<b200tag_blob>(hd, B200TAG_STAGE_CLUSTERS);
  if (s.pairs.size() > 4096) Die("more than 4096 blob pairs: do not fit the reference's 12-bit blob index (points.h:183-192)");
  std::sort(s.pairs.begin(), s.pairs.end(),
            [](const b200tag_blob &a, const b200tag_blob &b) { return PairKey(a.rep0, a.rep1) < PairKey(b.rep0, b.rep1); });
  std::unordered_map<uint32_t, uint32_t> pair_of_slot;
  std::unordered_map<uint64_t, uint32_t> pair_of_key;
  s.first_point.resize(s.pairs.size());
  uint32_t run = 0;
  for (uint32_t i = 0; i < s.pairs.size(); i++) {
    pair_of_slot[s.pairs[i].slot] = i;
    pair_of_key[PairKey(s.pairs[i].rep0, s.pairs[i].rep1)] = i;
    s.first_point[i] = run;
    run += s.pairs[i].count;
  }
  if (points) {
    const std::vector<b200tag_point> raw = Stage<b200tag_point>(hd, B200TAG_STAGE_POINTS);
    struct P {
      uint32_t pair, dense;
      QuadBoundaryPoint q;
    };
    std::vector<P> ps;
    ps.reserve(raw.size());
    const uint32_t plane = static_cast<uint32_t>((w - 2) * (h - 2));
    for (const b200tag_point &r : raw) {
      const auto it = pair_of_slot.find(r.slot);
      if (it == pair_of_slot.end()) Die("boundary point of an unknown blob pair");
      const int dx = r.dir == 2 ? 0 : (r.dir == 3 ? -1 : 1), dy = r.dir == 0 ? 0 : 1;
      const uint32_t bx = (r.x - dx) / 2, by = (r.y - dy) / 2;
      P p;
      p.pair = it->second;
      p.dense = plane * r.dir + (bx - 1) + (by - 1) * static_cast<uint32_t>(w - 2);
      p.q.set_rep0(s.pairs[p.pair].rep0);
      p.q.set_rep1(s.pairs[p.pair].rep1);
      p.q.set_base_xy(bx, by);
      p.q.set_dxy(r.dir);
      p.q.set_black_to_white(r.black_to_white != 0);
      ps.push_back(p);
    }
    std::sort(ps.begin(), ps.end(), [](const P &a, const P &b) { return a.dense < b.dense; });
    s.compressed.reserve(ps.size());
    for (const P &p : ps) s.compressed.push_back(p.q);
    std::stable_sort(ps.begin(), ps.end(), [](const P &a, const P &b) { return a.pair < b.pair; });
    s.sorted.reserve(ps.size());
    for (const P &p : ps) s.sorted.push_back(p.q);
  }
  if (per_blob) {
    s.blobs = Stage<b200tag_blob>(hd, B200TAG_STAGE_BLOBS);
    s.keys = Stage<uint64_t>(hd, B200TAG_STAGE_SORTED_POINTS);
    s.lfp = Stage<b200tag_lfp>(hd, B200TAG_STAGE_LINE_FIT_POINTS);
    s.errs = Stage<float>(hd, B200TAG_STAGE_ERRORS);
    s.filt = Stage<double>(hd, B200TAG_STAGE_FILTERED_ERRORS);
    s.pair_of_blob.assign(s.blobs.size(), 0);
    std::vector<int> blob_of_pair(s.pairs.size(), -1);
    for (uint32_t b = 0; b < s.blobs.size(); b++) {
      const auto it = pair_of_key.find(PairKey(s.blobs[b].rep0, s.blobs[b].rep1));
      if (it == pair_of_key.end()) Die("candidate blob of an unknown blob pair");
      s.pair_of_blob[b] = it->second;
      blob_of_pair[it->second] = static_cast<int>(b);
    }
    uint32_t first = 0;
    for (uint32_t i = 0; i < s.pairs.size(); i++) {
      if (!s.pairs[i].selected) continue;
      if (blob_of_pair[i] < 0) Die("selected blob pair without a candidate blob");
      const b200tag_blob &b = s.blobs[blob_of_pair[i]];
      s.sel.push_back(Snapshot::Sel{i, b.offset, b.count, first});
      first += b.count;
    }
  }
  return s;
}

IndexPoint ToIndexPoint(uint32_t pair, uint64_t key, bool black_to_white) {
  QuadBoundaryPoint q;
  q.set_base_xy(static_cast<uint32_t>(key) & 0xfff, static_cast<uint32_t>(key >> 12) & 0xfff);
  q.set_dxy((key >> 24) & 3);
  q.set_black_to_white(black_to_white);
  IndexPoint ip(pair, q.point_bits());
  ip.set_theta(static_cast<uint32_t>(key >> 26) & 0xfffffff);
  return ip;
}

// the strict local maxima of the filtered errors of one blob (line_fit_filter.cu:582), cyclic
bool IsPeak(const double *f, uint32_t n, uint32_t i) { return f[i] > f[(i + 1) % n] && f[i] > f[(i + n - 1) % n]; }

}  // namespace

#define QUAD_DIMS                                                          \
  const int f_ = static_cast<int>(tag_detector_->quad_decimate);           \
  const int w_ = static_cast<int>(width_) / f_, h_ = static_cast<int>(height_) / f_

void GpuDetector::CopyUnionMarkerPairTo(QuadBoundaryPoint *output) const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, true, false);
  const size_t total = static_cast<size_t>(4) * (w_ - 2) * (h_ - 2);
  std::fill(output, output + total, QuadBoundaryPoint());
  const uint32_t plane = static_cast<uint32_t>((w_ - 2) * (h_ - 2));
  for (const QuadBoundaryPoint &q : s.compressed)
    output[plane * (q.key & 3) + (q.base_x() - 1) + (q.base_y() - 1) * static_cast<uint32_t>(w_ - 2)] = q;
}

void GpuDetector::CopyCompressedUnionMarkerPairTo(QuadBoundaryPoint *output) const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, true, false);
  std::copy(s.compressed.begin(), s.compressed.end(), output);
}

std::vector<QuadBoundaryPoint> GpuDetector::CopySortedUnionMarkerPair() const {
  QUAD_DIMS;
  return Take(*this, w_, h_, true, false).sorted;
}

std::vector<MinMaxExtents> GpuDetector::CopyExtents() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, false);
  std::vector<MinMaxExtents> out(s.pairs.size());
  for (size_t i = 0; i < s.pairs.size(); i++) {
    const b200tag_blob &b = s.pairs[i];
    out[i] = MinMaxExtents{static_cast<uint16_t>(b.min_x), static_cast<uint16_t>(b.min_y), static_cast<uint16_t>(b.max_x),
                           static_cast<uint16_t>(b.max_y), s.first_point[i], b.count, b.gx_sum, b.gy_sum, b.pxgx_plus_pygy_sum};
  }
  return out;
}

std::vector<cub::KeyValuePair<long, MinMaxExtents>> GpuDetector::CopySelectedExtents() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, false);
  // TransformZeroFilteredBlobSizes + the SumPoints scan (apriltag_gpu.cu:582-629,873-905): rejected pairs keep their
  // box with count 0; starting_offset counts the points of the selected pairs before this one
  std::vector<cub::KeyValuePair<long, MinMaxExtents>> out(s.pairs.size());
  uint32_t before = 0;
  for (size_t i = 0; i < s.pairs.size(); i++) {
    const b200tag_blob &b = s.pairs[i];
    const uint32_t count = b.selected ? b.count : 0;
    out[i].key = static_cast<long>(i);
    out[i].value = MinMaxExtents{static_cast<uint16_t>(b.min_x), static_cast<uint16_t>(b.min_y), static_cast<uint16_t>(b.max_x),
                                 static_cast<uint16_t>(b.max_y), before, count, 0, 0, 0};
    before += count;
  }
  return out;
}

std::vector<IndexPoint> GpuDetector::CopySortedSelectedBlobs() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, true, true);
  std::vector<IndexPoint> out;
  for (const Snapshot::Sel &b : s.sel) {
    // black_to_white is not part of the engine's sort key: take it from the pair's boundary points
    std::unordered_map<uint32_t, bool> b2w;
    for (uint32_t i = 0; i < s.pairs[b.pair].count; i++) {
      const QuadBoundaryPoint &q = s.sorted[s.first_point[b.pair] + i];
      b2w[q.point_bits() & ~8u] = q.black_to_white();
    }
    for (uint32_t i = 0; i < b.count; i++) {
      IndexPoint ip = ToIndexPoint(b.pair, s.keys[b.seg + i], false);
      if (b2w[ip.point_bits() & ~8u]) ip.key |= 8;
      out.push_back(ip);
    }
  }
  return out;
}

std::vector<IndexPoint> GpuDetector::CopySelectedBlobs() const {
  // before the angle sort the points of a blob are in the order of the sorted boundary list: (dir, y, x)
  std::vector<IndexPoint> out = CopySortedSelectedBlobs();
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  for (const Snapshot::Sel &b : s.sel)
    std::sort(out.begin() + b.first, out.begin() + b.first + b.count, [](const IndexPoint &a, const IndexPoint &c) {
      const uint64_t ka = (static_cast<uint64_t>(a.key & 3) << 20) | (static_cast<uint64_t>(a.base_y()) << 10) | a.base_x();
      const uint64_t kc = (static_cast<uint64_t>(c.key & 3) << 20) | (static_cast<uint64_t>(c.base_y()) << 10) | c.base_x();
      return ka < kc;
    });
  return out;
}

std::vector<LineFitPoint> GpuDetector::CopyLineFitPoints() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  std::vector<LineFitPoint> out;
  for (const Snapshot::Sel &b : s.sel)
    for (uint32_t i = 0; i < b.count; i++) {
      const b200tag_lfp &l = s.lfp[b.seg + i];
      out.push_back(LineFitPoint{l.Mxx, l.Myy, l.Mxy, static_cast<int32_t>(l.Mx), static_cast<int32_t>(l.My), static_cast<int32_t>(l.W), b.pair});
    }
  return out;
}

std::vector<double> GpuDetector::CopyErrors() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  std::vector<double> out;
  for (const Snapshot::Sel &b : s.sel)
    for (uint32_t i = 0; i < b.count; i++) out.push_back(static_cast<double>(s.errs[b.seg + i]));
  return out;
}

std::vector<double> GpuDetector::CopyFilteredErrors() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  std::vector<double> out;
  for (const Snapshot::Sel &b : s.sel) out.insert(out.end(), s.filt.begin() + b.seg, s.filt.begin() + b.seg + b.count);
  return out;
}

std::vector<Peak> GpuDetector::CopyPeaks() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  std::vector<Peak> out;
  for (const Snapshot::Sel &b : s.sel)
    for (uint32_t i = 0; i < b.count; i++) {
      const bool peak = IsPeak(s.filt.data() + b.seg, b.count, i);
      out.push_back(Peak{static_cast<float>(-s.filt[b.seg + i]), b.first + i, peak ? static_cast<uint16_t>(b.pair) : Peak::kNoPeak()});
    }
  return out;
}

std::vector<Peak> GpuDetector::CopyCompressedPeaks() const {
  std::vector<Peak> all = CopyPeaks(), out;
  for (const Peak &p : all)
    if (p.blob_index != Peak::kNoPeak()) out.push_back(p);
  // C9, apriltag_gpu.cu:1017-1034: by blob, then by error (minus the filtered error: strongest first)
  std::stable_sort(out.begin(), out.end(), [](const Peak &a, const Peak &b) {
    return a.blob_index != b.blob_index ? a.blob_index < b.blob_index : a.error < b.error;
  });
  return out;
}

int GpuDetector::NumCompressedPeaks() const { return static_cast<int>(CopyCompressedPeaks().size()); }

std::vector<FitQuad> GpuDetector::CopyFitQuads() const {
  QUAD_DIMS;
  const Snapshot s = Take(*this, w_, h_, false, true);
  const std::vector<b200tag_fit_quad> fq = Stage<b200tag_fit_quad>(handle(), B200TAG_STAGE_FIT_QUADS);
  std::vector<FitQuad> out;
  for (const b200tag_fit_quad &q : fq) {
    FitQuad o;
    o.blob_index = static_cast<uint16_t>(s.pair_of_blob[q.blob_index]);
    o.valid = q.valid != 0;
    for (int i = 0; i < 4; i++) {
      o.indices[i] = static_cast<uint16_t>(q.indices[i]);
      const b200tag_moments &m = q.moments[i];
      o.moments[i] = LineFitMoments{static_cast<int32_t>(m.Mx), static_cast<int32_t>(m.My), static_cast<int32_t>(m.W), m.Mxx, m.Myy, m.Mxy, m.N};
    }
    out.push_back(o);
  }
  std::sort(out.begin(), out.end(), [](const FitQuad &a, const FitQuad &b) { return a.blob_index < b.blob_index; });
  return out;
}

}  // namespace frc971::apriltag
