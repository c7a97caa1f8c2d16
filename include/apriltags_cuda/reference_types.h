// The reference's packed device-side record types, as HOST types, for the typed debug accessors of
// frc971::apriltag::GpuDetector (reference: apriltag_gpu.h:111-183; the records themselves: points.h:25-279,
// line_fit_filter.h:14-135).  Field meaning, bit positions and struct layouts follow the reference so that code written
// against its accessors keeps compiling and reading the same bits; the engine itself never uses these types (its own
// records are wider: 12-bit coordinates, 20-bit slots -- csrc/dev_types.h), the accessors convert on the host.
//
// Limits inherited from the reference's bit fields: quad image at most 1024 x 1024 (10-bit base coordinates), labels
// below 2^20 (20-bit blob ids), at most 4096 blob pairs (12-bit blob index).  A frame that does not fit makes the typed
// accessors abort with a message; b200tag_copy_stage() has no such limits.
#ifndef B200TAG_APRILTAGS_CUDA_REFERENCE_TYPES_H_
#define B200TAG_APRILTAGS_CUDA_REFERENCE_TYPES_H_

#include <cstddef>
#include <cstdint>

#if defined(__has_include)
#if __has_include(<cub/util_type.cuh>) && defined(__CUDACC__)
#include <cub/util_type.cuh>
#define B200TAG_HAVE_CUB_KVP 1
#endif
#endif
#ifndef B200TAG_HAVE_CUB_KVP
namespace cub {  // host builds without CUB: the one template the accessor signatures name
template <typename K, typename V>
struct KeyValuePair {
  K key;
  V value;
};
}  // namespace cub
#endif

namespace frc971::apriltag {

namespace reference_bits {
inline int32_t DirDx(uint64_t key) { return (key & 3) == 2 ? 0 : ((key & 3) == 3 ? -1 : 1); }
inline int32_t DirDy(uint64_t key) { return (key & 3) == 0 ? 0 : 1; }
}  // namespace reference_bits

// points.h:25-161.  key = rep1[63:44] | rep0[43:24] | base x[23:14] | base y[13:4] | black_to_white[3] | dir[1:0]
struct QuadBoundaryPoint {
  uint64_t key = 0;

  uint32_t rep0() const { return (key >> 24) & 0xfffff; }
  uint32_t rep1() const { return (key >> 44) & 0xfffff; }
  uint64_t rep01() const { return (key >> 24) & 0xffffffffffull; }
  uint32_t point_bits() const { return key & 0xffffff; }
  uint32_t base_x() const { return (key >> 14) & 0x3ff; }
  uint32_t base_y() const { return (key >> 4) & 0x3ff; }
  int32_t dx() const { return reference_bits::DirDx(key); }
  int32_t dy() const { return reference_bits::DirDy(key); }
  uint32_t x() const { return static_cast<int32_t>(base_x() * 2) + dx(); }
  uint32_t y() const { return static_cast<int32_t>(base_y() * 2) + dy(); }
  bool black_to_white() const { return (key & 8) != 0; }
  int8_t gx() const { return black_to_white() ? dx() : -dx(); }
  int8_t gy() const { return black_to_white() ? dy() : -dy(); }
  bool nonzero() const { return key != 0; }

  void set_rep0(uint32_t v) { key = (key & 0xfffff00000ffffffull) | (static_cast<uint64_t>(v & 0xfffff) << 24); }
  void set_rep1(uint32_t v) { key = (key & 0x00000fffffffffffull) | (static_cast<uint64_t>(v & 0xfffff) << 44); }
  void set_base_xy(uint32_t x, uint32_t y) {
    key = (key & 0xffffffffff00000full) | (static_cast<uint64_t>(x & 0x3ff) << 14) | (static_cast<uint64_t>(y & 0x3ff) << 4);
  }
  void set_dxy(uint64_t d) { key = (key & ~3ull) | (d & 3); }
  void set_black_to_white(bool b) { key = (key & ~8ull) | (static_cast<uint64_t>(b) << 3); }

  bool operator==(const QuadBoundaryPoint &o) const { return key == o.key; }
  bool operator!=(const QuadBoundaryPoint &o) const { return key != o.key; }
  bool operator<(const QuadBoundaryPoint &o) const { return key < o.key; }
};

// points.h:169-279.  key = blob index[63:52] | theta[51:24] | the 24 point bits of a QuadBoundaryPoint
struct IndexPoint {
  static constexpr size_t kMaxBlobs = 2048;
  uint64_t key = 0;

  IndexPoint() = default;
  IndexPoint(uint32_t blob_index, uint32_t point_bits)
      : key((static_cast<uint64_t>(blob_index & 0xfff) << 52) | static_cast<uint64_t>(point_bits & 0xffffff)) {}

  uint32_t blob_index() const { return (key >> 52) & 0xfff; }
  uint32_t theta() const { return (key >> 24) & 0xfffffff; }
  uint32_t point_bits() const { return key & 0xffffff; }
  uint32_t base_x() const { return (key >> 14) & 0x3ff; }
  uint32_t base_y() const { return (key >> 4) & 0x3ff; }
  int32_t dx() const { return reference_bits::DirDx(key); }
  int32_t dy() const { return reference_bits::DirDy(key); }
  uint32_t x() const { return static_cast<int32_t>(base_x() * 2) + dx(); }
  uint32_t y() const { return static_cast<int32_t>(base_y() * 2) + dy(); }
  bool black_to_white() const { return (key & 8) != 0; }
  int8_t gx() const { return black_to_white() ? dx() : -dx(); }
  int8_t gy() const { return black_to_white() ? dy() : -dy(); }

  void set_blob_index(uint32_t v) { key = (key & 0x000fffffffffffffull) | (static_cast<uint64_t>(v & 0xfff) << 52); }
  void set_theta(uint32_t v) { key = (key & 0xfff0000000ffffffull) | (static_cast<uint64_t>(v & 0xfffffff) << 24); }
};

// line_fit_filter.h:14-59 (coordinates in half-pixel units of the quad image)
struct MinMaxExtents {
  uint16_t min_x, min_y, max_x, max_y;
  uint32_t starting_offset;
  uint32_t count;
  int32_t gx_sum, gy_sum;
  int64_t pxgx_plus_pygy_sum;

  double cx() const { return (min_x + max_x) * 0.5f + 0.05118; }
  double cy() const { return (min_y + max_y) * 0.5f + -0.028581; }
  float dot() const {
    return static_cast<double>(pxgx_plus_pygy_sum * 2 - (min_x + max_x) * gx_sum - (min_y + max_y) * gy_sum) * 0.5 -
           0.05118 * static_cast<double>(gx_sum) + 0.028581 * static_cast<double>(gy_sum);
  }
};

// line_fit_filter.h:61-83: inclusive per-blob prefix moments
struct alignas(16) LineFitPoint {
  int64_t Mxx, Myy, Mxy;
  int32_t Mx, My, W;
  uint32_t blob_index;
};

// line_fit_filter.h:85-94
struct LineFitMoments {
  int32_t Mx, My, W;
  int64_t Mxx, Myy, Mxy;
  int N;
};

// line_fit_filter.h:99-106
struct Peak {
  static constexpr uint16_t kNoPeak() { return 0xffff; }
  float error;                    // minus the filtered line-fit error: ascending = strongest first
  uint32_t filtered_point_index;  // index among the selected, sorted points
  uint16_t blob_index;            // kNoPeak() if this point is no local maximum
};

// line_fit_filter.h:130-135
struct FitQuad {
  uint16_t blob_index;
  bool valid;
  uint16_t indices[4];
  LineFitMoments moments[4];
};

}  // namespace frc971::apriltag

#endif
