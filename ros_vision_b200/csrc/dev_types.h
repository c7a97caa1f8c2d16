// Device-side data layout of the B200 AprilTag engine (shared by all .cu files).
//
// One `FrameParams` describes the HBM layout of a batch: every per-frame array is
// `base + frame * stride`; kernels take the struct by value and use blockIdx.y
// (or a device-side loop) as the frame index.  See DESIGN.md "Data layout in HBM".
#ifndef B200TAG_DEV_TYPES_H_
#define B200TAG_DEV_TYPES_H_

#include <stdint.h>

#include "../../include/b200tag.h"

namespace b200tag {

constexpr uint32_t kMinBlobPixels = 25;      // apriltag_gpu.cu:284,306 (union_markers_size >= 25)
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kMaxPeaks = 10;                // line_fit_filter.cu:630 (kNMaxima)
constexpr int kNumCombos = 210;              // line_fit_filter.h:160
constexpr uint32_t kSmallBlobPoints = 192;   // blobs up to this size are fitted by a single warp

// Packed boundary point, 64 bit:
//   [62:43] cluster slot | [42:27] rank of the point inside its cluster (saturating) | [26:15] base x |
//   [14:3] base y | [2:1] dir | [0] black_to_white
// (the reference's QuadBoundaryPoint, points.h:25-161, carries 20-bit blob ids and 10-bit
// coordinates; the slot indirection and 12-bit base coordinates lift its 1024x1024 quad-image
// limit to 4096x4096).  The rank is handed out while the point is counted, so the scatter into
// per-blob segments needs no second atomic; clusters with more than 65535 points are far above
// max_cluster_pixels (<= 4 * (4096 + 4096)) and never selected, so saturation is harmless.
// The low 27 bits are what travels on into the blob segments ("segment point").
constexpr uint32_t kPointRankMax = 0xffffu;
constexpr uint32_t kInvalidSlot = 0xfffffu;
__host__ __device__ inline uint64_t pack_point(uint32_t slot, uint32_t rank, uint32_t bx, uint32_t by, uint32_t dir, uint32_t b2w) {
  return (static_cast<uint64_t>(slot) << 43) | (static_cast<uint64_t>(rank < kPointRankMax ? rank : kPointRankMax) << 27) |
         (static_cast<uint64_t>(bx) << 15) | (static_cast<uint64_t>(by) << 3) | (static_cast<uint64_t>(dir) << 1) | b2w;
}
__host__ __device__ inline int dir_dx(uint32_t d) { return d == 2 ? 0 : (d == 3 ? -1 : 1); }
__host__ __device__ inline int dir_dy(uint32_t d) { return d == 0 ? 0 : 1; }
__host__ __device__ inline uint32_t point_slot(uint64_t p) { return static_cast<uint32_t>(p >> 43) & 0xfffffu; }
__host__ __device__ inline uint32_t point_rank(uint64_t p) { return static_cast<uint32_t>(p >> 27) & 0xffffu; }
__host__ __device__ inline uint32_t point_seg(uint64_t p) { return static_cast<uint32_t>(p) & 0x7ffffffu; }
__host__ __device__ inline uint32_t sp_bx(uint32_t s) { return (s >> 15) & 0xfffu; }
__host__ __device__ inline uint32_t sp_by(uint32_t s) { return (s >> 3) & 0xfffu; }
__host__ __device__ inline uint32_t sp_dir(uint32_t s) { return (s >> 1) & 3u; }
__host__ __device__ inline uint32_t sp_b2w(uint32_t s) { return s & 1u; }
// half-pixel coordinates, points.h:111-116
__host__ __device__ inline uint32_t sp_x(uint32_t s) { return 2 * sp_bx(s) + dir_dx(sp_dir(s)); }
__host__ __device__ inline uint32_t sp_y(uint32_t s) { return 2 * sp_by(s) + dir_dy(sp_dir(s)); }

// Sort key of a selected point inside its blob, 64 bit:
//   [53:26] theta (28 bit, IndexPoint::theta, points.h:195-202) | [25:24] dir | [23:12] base y | [11:0] base x
// Ascending key order == the reference's stable radix sort on (blob, theta) over the
// (dir, y, x)-ordered compaction (apriltag_gpu.cu:788-825,944-956).
__host__ __device__ inline uint64_t pack_sort_key(uint32_t theta, uint32_t dir, uint32_t by, uint32_t bx) {
  return (static_cast<uint64_t>(theta) << 26) | (static_cast<uint64_t>(dir) << 24) | (static_cast<uint64_t>(by) << 12) | bx;
}
__host__ __device__ inline uint32_t key_theta(uint64_t k) { return static_cast<uint32_t>(k >> 26) & 0xfffffff; }
__host__ __device__ inline uint32_t key_dir(uint64_t k) { return static_cast<uint32_t>(k >> 24) & 3; }
__host__ __device__ inline uint32_t key_by(uint64_t k) { return static_cast<uint32_t>(k >> 12) & 0xfff; }
__host__ __device__ inline uint32_t key_bx(uint64_t k) { return static_cast<uint32_t>(k) & 0xfff; }

struct Counters {  // one per frame, zeroed before each frame
  uint32_t status;
  uint32_t num_points;
  unsigned long long alloc;  // [63:40] blob pairs | [39:20] candidate blobs | [19:0] small blobs (k_select)
  uint32_t num_selected_points;
  uint32_t num_fit_quads;
  uint32_t num_quads;
  uint32_t num_detections;
  uint32_t next_quad;   // dynamic work counters
  uint32_t next_small;
  uint32_t next_large;
  uint32_t num_occupied;  // claimed hash slots (listed in FrameParams::occupied)
  uint32_t next_medium;
  uint32_t num_seg_points;      // points of all candidate blobs (count filter only): segment allocation
  uint32_t num_selected_blobs;  // candidates that also pass SelectBlobs' extent / polarity tests
  uint32_t num_medium;          // candidates of the medium tier (front of large_list)
  uint32_t num_large;           // candidates of the large tier (back of large_list)
  uint32_t num_huge;            // candidates above the large tier's shared-memory capacity (back of small_list)
  uint32_t next_huge;
  uint32_t pad[13];
};
__host__ __device__ inline uint32_t alloc_clusters(unsigned long long a) { return static_cast<uint32_t>(a >> 40); }
__host__ __device__ inline uint32_t alloc_blobs(unsigned long long a) { return static_cast<uint32_t>(a >> 20) & 0xfffffu; }
__host__ __device__ inline uint32_t alloc_small(unsigned long long a) { return static_cast<uint32_t>(a) & 0xfffffu; }
static_assert(sizeof(Counters) == 128, "Counters layout");

// What the fit kernels hand to k_quads for every blob with at least one peak: the (at most 10) strongest
// peaks in position order and exactly the prefix-moment records ReadMoments (line_fit_filter.cu:745-796)
// can touch for ranges between them -- lf[idx], lf[idx - 1] and lf[cnt - 1].
struct alignas(16) PeakTable {  // 1072 bytes: a multiple of 16, moved with 16-byte accesses
  uint32_t blob, cnt, nsel, npk;
  uint32_t rep0, rep1;
  uint32_t idx[kMaxPeaks];
  b200tag_lfp at[kMaxPeaks];      // lf[idx[k]]
  b200tag_lfp before[kMaxPeaks];  // lf[idx[k] - 1] (zero when idx[k] == 0)
  b200tag_lfp last;               // lf[cnt - 1]
};
static_assert(sizeof(PeakTable) % 16 == 0, "PeakTable is copied in 16-byte pieces");

// One entry of the fit kernels' work lists: everything a kernel needs to start loading the blob's points, so that the
// blob record itself is off the critical path.  blob = 0xffffffff marks a slot whose blob did not fit the buffers.
struct alignas(16) WorkItem {
  uint32_t blob, offset, count, pad;
};

// A tag family on the device (apriltag_family_t, the fields quad_decode reads).
constexpr int kMaxFamilies = 8;
constexpr int kMaxFamilyBits = 64;
constexpr int kMaxTotalWidth = 12;
struct DevFamily {
  uint32_t nbits, ncodes;
  int32_t width_at_border, total_width;
  int32_t reversed_border, max_hamming;
  uint32_t codes_off;  // first code in FrameParams::family_codes
  uint32_t pad;
  int8_t bit_x[kMaxFamilyBits], bit_y[kMaxFamilyBits];
};

struct FrameParams {
  // geometry
  int32_t W, H;        // full resolution
  int32_t w, h;        // quad image
  int32_t f;           // quad_decimate
  int32_t fmt;
  int32_t tiles_x, tiles_y;  // 4x4 threshold tiles
  uint32_t inv_w;      // ceil(2^32 / w): x / w == __umulhi(x, inv_w) for x < 2^20
  int32_t blur_ksz;    // 0 = no blur
  uint8_t blur_k[32];
  int32_t sharpen;     // quad_sigma < 0
  // detector parameters
  int32_t min_white_black_diff;
  uint32_t min_cluster_pixels;  // max(24, qtp.min_cluster_pixels)
  uint32_t max_cluster_pixels;  // 4 * (w + h)
  int32_t min_tag_width;
  int32_t normal_border, reversed_border;
  float cos_critical_rad;
  float max_line_fit_mse;
  int32_t refine_edges;
  double decode_sharpening;
  double fx, cx, fy, cy, k1, k2, p1, p2, k3;
  int32_t keep_stages;
  int32_t test_flags;  // B200TAG_TEST_*
  // capacities
  uint32_t point_cap, hash_cap /* pow2 */, blob_cap, quad_cap, det_cap, cluster_cap;
  // buffers (frame 0) and per-frame strides in elements
  const uint8_t *in;    size_t in_stride;
  uint8_t *gray;        size_t gray_stride;      // W*H (aliases `in` for GRAY8)
  uint8_t *quad;                                  // w*h
  uint8_t *quad_tmp;                              // w*h, blur scratch (same stride)
  uint8_t *minmax_raw;                            // tiles*2
  uint8_t *minmax;                                // tiles*2 filtered (keep_stages)
  uint8_t *thresh;                                // w*h
  uint32_t *labels;                               // w*h label words: [27:0] label | [28] component >= 25 px | [30:29] colour
  uint32_t *sizes;                                // w*h
  uint32_t *tile_roots;   // per CCL tile: tile roots that touch the tile border (k_ccl_local -> k_ccl_handoff)
  uint32_t *tile_nroots;  // per CCL tile: entries in use
  uint64_t *points;     // point_cap
  // blob-pair hash (hash_cap each): key and point count; extents are computed per blob later
  unsigned long long *h_key;
  uint32_t *h_count;
  uint32_t *slot_off;   // hash_cap: first segment position of the slot's blob, 0xffffffff = not a candidate
  uint32_t *slot_cluster;  // hash_cap: index into clusters[] (keep_stages)
  uint32_t *occupied;   // hash_cap: slots claimed this frame, in claim order
  b200tag_blob *blobs;  // blob_cap
  uint32_t *seg_pts;    // point_cap: segment points (27 bit) of candidate blobs, unsorted
  WorkItem *small_list; // blob_cap: blobs fitted by one warp (front), by a 512-thread CTA (back)
  WorkItem *large_list; // blob_cap: blobs fitted by a 128-thread CTA (front), by a 256-thread CTA (back)
  b200tag_blob *clusters;  // cluster_cap (keep_stages)
  uint64_t *seg_keys;   // point_cap
  b200tag_lfp *lfp;     // point_cap
  float *errs;          // point_cap
  double *filt;         // point_cap
  uint64_t *peak_ws;    // point_cap / 2 + 1: peak list of blobs too large for shared memory
  PeakTable *peak_tables;       // blob_cap: hand-off from the fit kernels to k_quads
  b200tag_fit_quad *fit_quads;  // blob_cap
  b200tag_quad *quads;  // quad_cap
  b200tag_detection *dets;  // det_cap
  const DevFamily *families;      // nfamilies (device memory, owned by the detector)
  const uint64_t *family_codes;   // all families' code tables back to back
  int32_t nfamilies;
  Counters *counters;
};

}  // namespace b200tag

#endif
