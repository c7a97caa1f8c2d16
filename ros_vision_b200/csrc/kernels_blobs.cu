// Blob stage of the B200 AprilTag engine (sm_100a):
//   K7  blob selection straight off the blob-pair hash     (reference: apriltag_gpu.cu:522-629,873-905)
//   K8  scatter of surviving points into per-blob segments, keyed by angle (:380-412,909-942)
//   K9  one CTA per blob: in-CTA bitonic angle sort, prefix moments, line-fit errors, 7-tap
//       smoothing, peak selection, exhaustive 210-way quad search, corner/area/angle tests
//       (:944-1097, line_fit_filter.cu:22-36,217-278,504-592,709-1193, apriltag_detect.cu:38-282)
// The reference spends 10 CUB device-wide passes and 5 host round trips here; per-blob work is
// independent, so each blob is carried from unsorted points to QuadCorners by a single CTA with
// no host involvement.  Compiled with -fmad=false: the float/double expressions below are
// evaluated exactly as written (IEEE, no contraction), so a CPU restatement reproduces them bit for bit.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dev_types.h"
#include "kernels.h"

namespace b200tag {

// ---------------------------------------------------------------------------------------------
// K7
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool select_blob(const FrameParams &p, uint32_t count, uint32_t minx, uint32_t miny, uint32_t maxx,
                                            uint32_t maxy, int32_t gx, int32_t gy, long long dotsum) {
  // SelectBlobs::operator(), apriltag_gpu.cu:534-559
  if (count < p.min_cluster_pixels) return false;
  if (count > p.max_cluster_pixels) return false;
  if (static_cast<int>((maxx - minx) * (maxy - miny)) < p.min_tag_width) return false;
  // MinMaxExtents::dot(), line_fit_filter.h:51-58
  const long long a = dotsum * 2 - static_cast<long long>(static_cast<int>(minx + maxx) * gx) -
                      static_cast<long long>(static_cast<int>(miny + maxy) * gy);
  const double d = static_cast<double>(a) * 0.5 - 0.05118 * static_cast<double>(gx) + 0.028581 * static_cast<double>(gy);
  const bool quad_reversed = static_cast<float>(d) < 0.0;
  if (!p.reversed_border && quad_reversed) return false;
  if (!p.normal_border && !quad_reversed) return false;
  return true;
}

__global__ void __launch_bounds__(256) k_select(FrameParams p) {
  const int frame = blockIdx.y;
  const size_t hoff = static_cast<size_t>(frame) * p.hash_cap;
  Counters *ctr = p.counters + frame;
  b200tag_blob *blobs = p.blobs + static_cast<size_t>(frame) * p.blob_cap;
  uint32_t *fill = p.blob_fill + static_cast<size_t>(frame) * p.blob_cap;
  b200tag_blob *clusters = p.clusters ? p.clusters + static_cast<size_t>(frame) * p.cluster_cap : nullptr;
  const int lane = threadIdx.x & 31;
  // hash_cap is a multiple of the block size, so every warp runs the same trip count
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < p.hash_cap; slot += gridDim.x * blockDim.x) {
    const unsigned long long key = p.h_key[hoff + slot];
    const bool occ = key != kEmptyKey;
    b200tag_blob rec;
    bool sel = false;
    if (occ) {
      rec.rep0 = static_cast<uint32_t>(key >> 32);
      rec.rep1 = static_cast<uint32_t>(key);
      rec.min_x = p.h_minx[hoff + slot];
      rec.min_y = p.h_miny[hoff + slot];
      rec.max_x = p.h_maxx[hoff + slot];
      rec.max_y = p.h_maxy[hoff + slot];
      rec.count = p.h_count[hoff + slot];
      rec.gx_sum = p.h_gx[hoff + slot];
      rec.gy_sum = p.h_gy[hoff + slot];
      rec.pxgx_plus_pygy_sum = p.h_dot[hoff + slot];
      rec.slot = slot;
      rec.offset = 0;
      sel = select_blob(p, rec.count, rec.min_x, rec.min_y, rec.max_x, rec.max_y, rec.gx_sum, rec.gy_sum, rec.pxgx_plus_pygy_sum);
      rec.selected = sel;
      // leave the table empty for the next frame
      p.h_key[hoff + slot] = kEmptyKey;
      p.h_count[hoff + slot] = 0;
      p.h_minx[hoff + slot] = 0xffffffffu;
      p.h_miny[hoff + slot] = 0xffffffffu;
      p.h_maxx[hoff + slot] = 0;
      p.h_maxy[hoff + slot] = 0;
      p.h_gx[hoff + slot] = 0;
      p.h_gy[hoff + slot] = 0;
      p.h_dot[hoff + slot] = 0;
    }
    // warp-aggregated allocation: cluster index, blob index, point offset
    const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ);
    const uint32_t sel_mask = __ballot_sync(0xffffffffu, sel);
    uint32_t incl = sel ? rec.count : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t total_pts = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t cbase = 0, bbase = 0, pbase = 0;
    if (lane == 0) {
      if (occ_mask) cbase = atomicAdd(&ctr->num_clusters, __popc(occ_mask));
      if (sel_mask) {
        bbase = atomicAdd(&ctr->num_blobs, __popc(sel_mask));
        pbase = atomicAdd(&ctr->num_selected_points, total_pts);
      }
    }
    cbase = __shfl_sync(0xffffffffu, cbase, 0);
    bbase = __shfl_sync(0xffffffffu, bbase, 0);
    pbase = __shfl_sync(0xffffffffu, pbase, 0);
    int32_t blob_id = -1;
    if (sel) {
      const uint32_t b = bbase + __popc(sel_mask & ((1u << lane) - 1u));
      const uint32_t off = pbase + incl - rec.count;
      if (b < p.blob_cap && static_cast<uint64_t>(off) + rec.count <= p.point_cap) {
        rec.offset = off;
        blobs[b] = rec;
        fill[b] = 0;
        blob_id = static_cast<int32_t>(b);
      } else {
        atomicOr(&ctr->status, B200TAG_ST_BLOBS_OVERFLOW);
      }
    }
    if (occ && clusters) {
      const uint32_t c = cbase + __popc(occ_mask & ((1u << lane) - 1u));
      if (c < p.cluster_cap) clusters[c] = rec;
    }
    p.slot_blob[hoff + slot] = blob_id;
  }
}

// ---------------------------------------------------------------------------------------------
// K8
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(FrameParams p) {
  const int frame = blockIdx.y;
  const Counters *ctr = p.counters + frame;
  const uint32_t np = min(ctr->num_points, p.point_cap);
  const uint64_t *points = p.points + static_cast<size_t>(frame) * p.point_cap;
  const int32_t *slot_blob = p.slot_blob + static_cast<size_t>(frame) * p.hash_cap;
  const b200tag_blob *blobs = p.blobs + static_cast<size_t>(frame) * p.blob_cap;
  uint32_t *fill = p.blob_fill + static_cast<size_t>(frame) * p.blob_cap;
  uint64_t *seg = p.seg_keys + static_cast<size_t>(frame) * p.point_cap;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
    const uint64_t pt = points[i];
    const int32_t b = slot_blob[point_slot(pt)];
    if (b < 0) continue;  // NonzeroBlobs, apriltag_gpu.cu:505-518
    const uint32_t minx = blobs[b].min_x, maxx = blobs[b].max_x, miny = blobs[b].min_y, maxy = blobs[b].max_y;
    const uint32_t x = point_x(pt), y = point_y(pt), d = point_dir(pt);
    // MinMaxExtents::cx/cy, line_fit_filter.h:44-49
    const double cx = static_cast<double>(static_cast<float>(static_cast<int>(minx + maxx)) * 0.5f) + 0.05118;
    const double cy = static_cast<double>(static_cast<float>(static_cast<int>(miny + maxy)) * 0.5f) + -0.028581;
    // AddThetaToIndexPoint, apriltag_gpu.cu:400-408
    const float fy = static_cast<float>(static_cast<double>(y) - cy);
    const float fx = static_cast<float>(static_cast<double>(x) - cx);
    const float theta = static_cast<float>((static_cast<double>(atan2f(fy, fx)) + 3.14159265358979323846) * 8e6);
    long long ti = llrintf(theta);
    if (ti < 0) ti = 0;
    const uint32_t th = static_cast<uint32_t>(ti & 0xfffffff);
    const uint32_t bx = (x - dir_dx(d)) >> 1, by = (y - dir_dy(d)) >> 1;
    const uint32_t pos = blobs[b].offset + atomicAdd(&fill[b], 1u);
    seg[pos] = pack_sort_key(th, d, by, bx);
  }
}

// ---------------------------------------------------------------------------------------------
// K9
// ---------------------------------------------------------------------------------------------
constexpr int kFitThreads = 256;
constexpr int kSortCap = 4096;  // keys sorted in shared memory; larger blobs sort in place in L2

template <typename Ptr>
__device__ __forceinline__ void cmpxchg(Ptr a, uint32_t i, uint32_t l) {
  const uint64_t x = a[i], y = a[l];
  if (x > y) {
    a[i] = y;
    a[l] = x;
  }
}

// All-ascending bitonic network over N = pow2 >= cnt slots; slots >= cnt are virtual +inf and
// never move, so no padding is materialised.
template <typename Ptr>
__device__ void bitonic_sort(Ptr a, uint32_t cnt, uint32_t N) {
  for (uint32_t k = 2; k <= N; k <<= 1) {
    const uint32_t hk = k >> 1;
    for (uint32_t t = threadIdx.x; t < (N >> 1); t += kFitThreads) {
      const uint32_t blk = t / hk, w = t % hk;
      const uint32_t i = blk * k + w, l = blk * k + k - 1 - w;
      if (l < cnt) cmpxchg(a, i, l);
    }
    __syncthreads();
    for (uint32_t j = k >> 2; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < (N >> 1); t += kFitThreads) {
        const uint32_t i = 2 * j * (t / j) + (t % j), l = i + j;
        if (l < cnt) cmpxchg(a, i, l);
      }
      __syncthreads();
    }
  }
}

struct Mom {
  long long Mx, My, W, Mxx, Myy, Mxy;
  int N;
};

__device__ __forceinline__ b200tag_lfp load_lfp(const b200tag_lfp *p) {
  const longlong2 *q = reinterpret_cast<const longlong2 *>(p);
  const longlong2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2);
  b200tag_lfp r;
  r.Mxx = a.x; r.Myy = a.y; r.Mxy = b.x; r.Mx = b.y; r.My = c.x; r.W = c.y;
  return r;
}

// ReadMoments, line_fit_filter.cu:745-796 (== CalculateError's window logic, :230-274)
__device__ Mom read_moments(const b200tag_lfp *lf, uint32_t cnt, uint32_t i0, uint32_t i1) {
  Mom m;
  if (i0 < i1) {
    m.N = static_cast<int>(i1 - i0 + 1);
    const b200tag_lfp a = load_lfp(lf + i1);
    m.Mx = a.Mx; m.My = a.My; m.Mxx = a.Mxx; m.Mxy = a.Mxy; m.Myy = a.Myy; m.W = a.W;
    if (i0 > 0) {
      const b200tag_lfp b = load_lfp(lf + i0 - 1);
      m.Mx -= b.Mx; m.My -= b.My; m.Mxx -= b.Mxx; m.Mxy -= b.Mxy; m.Myy -= b.Myy; m.W -= b.W;
    }
  } else {
    const b200tag_lfp b = load_lfp(lf + i0 - 1), z = load_lfp(lf + cnt - 1), a = load_lfp(lf + i1);
    m.Mx = z.Mx - b.Mx + a.Mx;
    m.My = z.My - b.My + a.My;
    m.Mxx = z.Mxx - b.Mxx + a.Mxx;
    m.Mxy = z.Mxy - b.Mxy + a.Mxy;
    m.Myy = z.Myy - b.Myy + a.Myy;
    m.W = z.W - b.W + a.W;
    m.N = static_cast<int>(cnt - i0 + i1 + 1);
  }
  return m;
}

// FitLineError (line_fit_filter.cu:22-36) / FitLine (:798-872) / HostFitLine (apriltag_detect.cu:38-90)
__device__ __forceinline__ float eig_small_of(const Mom &m, float *hyp_out, long long *Cxx_o, long long *Cxy_o, long long *Cyy_o) {
  const long long Cxx = m.Mxx * m.W - m.Mx * m.Mx;
  const long long Cxy = m.Mxy * m.W - m.Mx * m.My;
  const long long Cyy = m.Myy * m.W - m.My * m.My;
  const float hyp = hypotf(static_cast<float>(Cxx - Cyy), static_cast<float>(2 * Cxy));
  const float eight_w2 = static_cast<float>(static_cast<double>(m.W * m.W) * 8.0);
  const float eig = (static_cast<float>(Cxx + Cyy) - hyp) / eight_w2;
  if (hyp_out) *hyp_out = hyp;
  if (Cxx_o) { *Cxx_o = Cxx; *Cxy_o = Cxy; *Cyy_o = Cyy; }
  return eig;
}

__device__ void fit_line(const Mom &m, double *lp01, double *lp23, double *err, double *mse) {
  float hyp;
  long long Cxx, Cxy, Cyy;
  const float eig = eig_small_of(m, &hyp, &Cxx, &Cxy, &Cyy);
  if (lp01) {
    lp01[0] = static_cast<double>(static_cast<float>(m.Mx) / static_cast<float>(m.W * 2));
    lp01[1] = static_cast<double>(static_cast<float>(m.My) / static_cast<float>(m.W * 2));
  }
  if (lp23) {
    const float nx1 = static_cast<float>(Cxx - Cyy) - hyp;
    const float ny1 = static_cast<float>(2 * Cxy);
    const float M1 = nx1 * nx1 + ny1 * ny1;
    const float nx2 = static_cast<float>(2 * Cxy);
    const float ny2 = static_cast<float>(Cyy - Cxx) - hyp;
    const float M2 = nx2 * nx2 + ny2 * ny2;
    float nx, ny;
    if (M1 > M2) { nx = nx1; ny = ny1; } else { nx = nx2; ny = ny2; }
    const float len = hypotf(nx, ny);
    lp23[0] = static_cast<double>(nx / len);
    lp23[1] = static_cast<double>(ny / len);
  }
  *err = static_cast<double>(static_cast<float>(m.N) * eig);
  *mse = static_cast<double>(eig);
}

// TransformLineFitPoint weight, apriltag_gpu.cu:644-657
__device__ __forceinline__ int point_weight(const uint8_t *im, int w, int h, int ix, int iy) {
  int W = 1;
  if (ix > 0 && ix + 1 < w && iy > 0 && iy + 1 < h) {
    const int gx = static_cast<int>(im[iy * w + ix + 1]) - static_cast<int>(im[iy * w + ix - 1]);
    const int gy = static_cast<int>(im[(iy + 1) * w + ix]) - static_cast<int>(im[(iy - 1) * w + ix]);
    W = static_cast<int>(hypotf(static_cast<float>(gx), static_cast<float>(gy)) + 1);
  }
  return W;
}

__device__ __forceinline__ uint32_t float_order(float f) {  // monotone float -> uint map (+0 == -0 after f + 0.0f)
  const uint32_t u = __float_as_uint(f + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr unsigned long long kNoKey = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, unsigned long long *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long r = threadIdx.x < (kFitThreads / 32) ? s_red[threadIdx.x] : kNoKey;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, r, o);
      r = t < r ? t : r;
    }
    if (threadIdx.x == 0) s_red[8] = r;
  }
  __syncthreads();
  const unsigned long long out = s_red[8];
  __syncthreads();
  return out;
}

struct FitShared {
  unsigned long long sort_buf[kSortCap];
  long long scan[6][kFitThreads];
  unsigned long long red[16];
  double cand_err[kFitThreads];
  uint32_t peak_idx[kMaxPeaks];
  uint8_t combos[kNumCombos][4];
  uint32_t cur_blob;
  uint32_t npeaks;
};

__global__ void __launch_bounds__(kFitThreads) k_fit_blobs(FrameParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FitShared &S = *reinterpret_cast<FitShared *>(smem_raw);
  const int frame = blockIdx.y;
  const int tid = threadIdx.x;
  Counters *ctr = p.counters + frame;
  const size_t n = static_cast<size_t>(p.w) * p.h;
  const uint8_t *quad = p.quad + frame * n;
  const b200tag_blob *blobs = p.blobs + static_cast<size_t>(frame) * p.blob_cap;
  uint64_t *seg_all = p.seg_keys + static_cast<size_t>(frame) * p.point_cap;
  b200tag_lfp *lfp_all = p.lfp + static_cast<size_t>(frame) * p.point_cap;
  float *errs_all = p.errs + static_cast<size_t>(frame) * p.point_cap;
  double *filt_all = p.filt + static_cast<size_t>(frame) * p.point_cap;
  b200tag_fit_quad *fit_quads = p.fit_quads + static_cast<size_t>(frame) * p.blob_cap;
  b200tag_quad *quads = p.quads + static_cast<size_t>(frame) * p.quad_cap;
  const uint32_t nblobs = min(ctr->num_blobs, p.blob_cap);

  if (tid == 0) {  // Unrank table: nested-loop order == line_fit_filter.cu:709-728
    int c = 0;
    for (int m0 = 0; m0 < kMaxPeaks - 3; m0++)
      for (int m1 = m0 + 1; m1 < kMaxPeaks - 2; m1++)
        for (int m2 = m1 + 1; m2 < kMaxPeaks - 1; m2++)
          for (int m3 = m2 + 1; m3 < kMaxPeaks; m3++) {
            S.combos[c][0] = m0; S.combos[c][1] = m1; S.combos[c][2] = m2; S.combos[c][3] = m3;
            c++;
          }
  }

  while (true) {
    __syncthreads();
    if (tid == 0) S.cur_blob = atomicAdd(&ctr->next_blob, 1u);
    __syncthreads();
    const uint32_t b = S.cur_blob;
    if (b >= nblobs) break;
    const b200tag_blob blob = blobs[b];
    const uint32_t cnt = blob.count, off = blob.offset;
    uint64_t *seg = seg_all + off;
    b200tag_lfp *lf = lfp_all + off;
    float *errs = errs_all + off;
    double *filt = filt_all + off;

    // (1) angle sort, C6 (apriltag_gpu.cu:944-956)
    uint32_t N = 1;
    while (N < cnt) N <<= 1;
    if (cnt <= kSortCap) {
      for (uint32_t i = tid; i < cnt; i += kFitThreads) S.sort_buf[i] = __ldcg(reinterpret_cast<const unsigned long long *>(seg + i));
      __syncthreads();
      bitonic_sort(S.sort_buf, cnt, N);
      for (uint32_t i = tid; i < cnt; i += kFitThreads) seg[i] = S.sort_buf[i];
    } else {
      bitonic_sort(reinterpret_cast<volatile unsigned long long *>(seg), cnt, N);
    }
    __syncthreads();

    // (2) weighted moments + inclusive prefix sums, C7 (apriltag_gpu.cu:631-687,984-987)
    const uint32_t chunk = (cnt + kFitThreads - 1) / kFitThreads;
    const uint32_t c_lo = min(cnt, tid * chunk), c_hi = min(cnt, c_lo + chunk);
    long long t_Mxx = 0, t_Myy = 0, t_Mxy = 0, t_Mx = 0, t_My = 0, t_W = 0;
    for (uint32_t i = c_lo; i < c_hi; i++) {
      const uint64_t k = seg[i];
      const uint32_t d = key_dir(k);
      const int ix2 = static_cast<int>(2 * key_bx(k)) + dir_dx(d) + 1, iy2 = static_cast<int>(2 * key_by(k)) + dir_dy(d) + 1;
      const long long W = point_weight(quad, p.w, p.h, ix2 / 2, iy2 / 2);
      t_Mx += W * ix2; t_My += W * iy2; t_Mxx += W * ix2 * ix2; t_Mxy += W * ix2 * iy2; t_Myy += W * iy2 * iy2; t_W += W;
    }
    S.scan[0][tid] = t_Mxx; S.scan[1][tid] = t_Myy; S.scan[2][tid] = t_Mxy;
    S.scan[3][tid] = t_Mx;  S.scan[4][tid] = t_My;  S.scan[5][tid] = t_W;
    __syncthreads();
    {  // exclusive scan of the 256 chunk totals: warp q scans quantity q
      const int warp = tid >> 5, lane = tid & 31;
      if (warp < 6) {
        long long v[8], run = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = S.scan[warp][lane * 8 + j]; run += v[j]; }
        long long incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        long long ex = incl - run;
#pragma unroll
        for (int j = 0; j < 8; j++) { S.scan[warp][lane * 8 + j] = ex; ex += v[j]; }
      }
    }
    __syncthreads();
    {
      long long a_Mxx = S.scan[0][tid], a_Myy = S.scan[1][tid], a_Mxy = S.scan[2][tid];
      long long a_Mx = S.scan[3][tid], a_My = S.scan[4][tid], a_W = S.scan[5][tid];
      for (uint32_t i = c_lo; i < c_hi; i++) {
        const uint64_t k = seg[i];
        const uint32_t d = key_dir(k);
        const int ix2 = static_cast<int>(2 * key_bx(k)) + dir_dx(d) + 1, iy2 = static_cast<int>(2 * key_by(k)) + dir_dy(d) + 1;
        const long long W = point_weight(quad, p.w, p.h, ix2 / 2, iy2 / 2);
        a_Mx += W * ix2; a_My += W * iy2; a_Mxx += W * ix2 * ix2; a_Mxy += W * ix2 * iy2; a_Myy += W * iy2 * iy2; a_W += W;
        longlong2 *q = reinterpret_cast<longlong2 *>(lf + i);
        q[0] = make_longlong2(a_Mxx, a_Myy);
        q[1] = make_longlong2(a_Mxy, a_Mx);
        q[2] = make_longlong2(a_My, a_W);
      }
    }
    __syncthreads();

    // (3) windowed line-fit error, K10 part 1 (line_fit_filter.cu:217-278)
    const uint32_t ksz = min(20u, cnt / 12u);
    for (uint32_t i = tid; i < cnt; i += kFitThreads) {
      const uint32_t i0 = (i + 2 * cnt - ksz) % cnt, i1 = (i + cnt + ksz) % cnt;
      const Mom m = read_moments(lf, cnt, i0, i1);
      const float eig = eig_small_of(m, nullptr, nullptr, nullptr, nullptr);
      errs[i] = static_cast<float>(m.N) * eig;
    }
    __syncthreads();
    // (4) 7-tap smoothing in double, (:504-525)
    for (uint32_t i = tid; i < cnt; i += kFitThreads) {
      const float kf[7] = {0.01110899634659290314f, 0.13533528149127960205f, 0.60653066635131835938f, 1.0f,
                           0.60653066635131835938f, 0.13533528149127960205f, 0.01110899634659290314f};
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const double e = static_cast<double>(__ldcg(errs + (i + cnt + j - 3) % cnt));
        acc += e * static_cast<double>(kf[j]);
      }
      filt[i] = acc;
    }
    __syncthreads();

    // (5) peaks: strict local maxima (:582); the 10 strongest by (-filtered as f32, index), C8-C10
    //     (apriltag_gpu.cu:1001-1078).  Ten rounds of "smallest key above the previous one".
    unsigned long long last = 0;
    uint32_t npk_local = 0;
    uint32_t nsel = 0;
    for (int round = 0; round < kMaxPeaks; round++) {
      unsigned long long best = kNoKey;
      for (uint32_t i = tid; i < cnt; i += kFitThreads) {
        const double m = __ldcg(filt + i);
        const double bv = __ldcg(filt + (i + cnt - 1) % cnt), av = __ldcg(filt + (i + 1) % cnt);
        if (m > bv && m > av) {
          if (round == 0) npk_local++;
          const unsigned long long key = (static_cast<unsigned long long>(float_order(static_cast<float>(-m))) << 32) | i;
          if ((round == 0 || key > last) && key < best) best = key;
        }
      }
      best = block_min_u64(best, S.red);
      if (best == kNoKey) break;
      if (tid == 0) S.peak_idx[round] = static_cast<uint32_t>(best & 0xffffffffu);
      last = best;
      nsel++;
    }
    {  // total number of peaks (PeakExtents.count)
      uint32_t v = npk_local;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __syncthreads();
      if ((tid & 31) == 0) S.red[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) {
        uint32_t s = 0;
        for (int q = 0; q < kFitThreads / 32; q++) s += static_cast<uint32_t>(S.red[q]);
        S.npeaks = s;
        // re-order the chosen maxima by position (line_fit_filter.cu:1104-1119)
        for (uint32_t a = 1; a < nsel; a++) {
          const uint32_t v2 = S.peak_idx[a];
          int q = static_cast<int>(a) - 1;
          while (q >= 0 && S.peak_idx[q] > v2) { S.peak_idx[q + 1] = S.peak_idx[q]; q--; }
          S.peak_idx[q + 1] = v2;
        }
      }
      __syncthreads();
    }
    if (S.npeaks == 0) continue;  // no PeakExtents entry -> no FitQuad

    // (6) exhaustive search over the C(10,4) corner choices, K11 (line_fit_filter.cu:889-1061,1088-1193)
    const int nm = static_cast<int>(nsel);
    double my_err = 1.7976931348623157e308;
    if (tid < kNumCombos) {
      const int m0 = S.combos[tid][0], m1 = S.combos[tid][1], m2 = S.combos[tid][2], m3 = S.combos[tid][3];
      if (m3 < nm) {
        const double max_mse = static_cast<double>(p.max_line_fit_mse);
        const double max_dot = static_cast<double>(p.cos_critical_rad);
        double e01, mse01, p01[2], e12, mse12, p12[2], e23, mse23, e30, mse30;
        bool ok = true;
        Mom mo = read_moments(lf, cnt, S.peak_idx[m0], S.peak_idx[m1]);
        fit_line(mo, nullptr, p01, &e01, &mse01);
        if (mse01 > max_mse) ok = false;
        if (ok) {
          mo = read_moments(lf, cnt, S.peak_idx[m1], S.peak_idx[m2]);
          fit_line(mo, nullptr, p12, &e12, &mse12);
          if (mse12 > max_mse) ok = false;
        }
        if (ok) {
          const double dot = p01[0] * p12[0] + p01[1] * p12[1];
          if (fabs(dot) > max_dot) ok = false;
        }
        if (ok) {
          mo = read_moments(lf, cnt, S.peak_idx[m2], S.peak_idx[m3]);
          fit_line(mo, nullptr, nullptr, &e23, &mse23);
          if (mse23 > max_mse) ok = false;
        }
        if (ok) {
          mo = read_moments(lf, cnt, S.peak_idx[m3], S.peak_idx[m0]);
          fit_line(mo, nullptr, nullptr, &e30, &mse30);
          if (mse30 > max_mse) ok = false;
        }
        if (ok) my_err = e01 + e12 + e23 + e30;
      }
    }
    S.cand_err[tid] = my_err;
    __syncthreads();

    // (7) finalisation by one thread: FitQuad, then UpdateFitQuads + AdjustPixelCenters
    //     (apriltag_detect.cu:98-282)
    if (tid == 0) {
      double best = 1.7976931348623157e308;
      int bi = 0;
      for (int c = 0; c < kNumCombos; c++)
        if (S.cand_err[c] < best) { best = S.cand_err[c]; bi = c; }  // ties keep the lowest rank
      const bool valid = best < static_cast<double>(p.max_line_fit_mse * static_cast<float>(cnt));
      const uint32_t fq = atomicAdd(&ctr->num_fit_quads, 1u);
      Mom moms[4];
      uint32_t idx[4] = {0, 0, 0, 0};
      if (valid) {
        for (int i = 0; i < 4; i++) idx[i] = S.peak_idx[S.combos[bi][i]];
        for (int i = 0; i < 4; i++) moms[i] = read_moments(lf, cnt, idx[i], idx[(i + 1) & 3]);
      }
      if (fq < p.blob_cap) {
        b200tag_fit_quad &o = fit_quads[fq];
        o.blob_index = b;
        o.valid = valid;
        o.num_peaks = static_cast<int32_t>(S.npeaks);
        o.err = best;
        for (int i = 0; i < 4; i++) {
          o.indices[i] = idx[i];
          if (valid) {
            o.moments[i].Mx = moms[i].Mx; o.moments[i].My = moms[i].My; o.moments[i].W = moms[i].W;
            o.moments[i].Mxx = moms[i].Mxx; o.moments[i].Myy = moms[i].Myy; o.moments[i].Mxy = moms[i].Mxy;
            o.moments[i].N = moms[i].N; o.moments[i].pad = 0;
          } else {
            o.moments[i] = b200tag_moments{0, 0, 0, 0, 0, 0, 0, 0};
          }
        }
      }
      if (valid) {
        double lines[4][4];
        for (int i = 0; i < 4; i++) {
          double err, mse;
          fit_line(moms[i], lines[i], lines[i] + 2, &err, &mse);
        }
        float cr[4][2];
        bool bad = false;
        for (int i = 0; i < 4 && !bad; i++) {  // apriltag_detect.cu:125-166
          const double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
          const double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
          const double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
          const double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
          const double det = A00 * A11 - A10 * A01;
          const double W00 = A11 / det, W01 = -A01 / det;
          if (fabs(det) < 0.001) { bad = true; break; }
          const double L0 = W00 * B0 + W01 * B1;
          cr[i][0] = static_cast<float>(lines[i][0] + L0 * A00);
          cr[i][1] = static_cast<float>(lines[i][1] + L0 * A10);
        }
        if (!bad) {  // :171-207
          float area = 0;
          float length[3], pp;
          for (int i = 0; i < 3; i++) {
            const int a = i, c = (i + 1) % 3;
            length[i] = hypotf(cr[c][0] - cr[a][0], cr[c][1] - cr[a][1]);
          }
          pp = (length[0] + length[1] + length[2]) / 2;
          area += sqrtf(pp * (pp - length[0]) * (pp - length[1]) * (pp - length[2]));
          const int idxs[4] = {2, 3, 0, 2};
          for (int i = 0; i < 3; i++) {
            const int a = idxs[i], c = idxs[i + 1];
            length[i] = hypotf(cr[c][0] - cr[a][0], cr[c][1] - cr[a][1]);
          }
          pp = (length[0] + length[1] + length[2]) / 2;
          area += sqrtf(pp * (pp - length[0]) * (pp - length[1]) * (pp - length[2]));
          if (static_cast<double>(area) < 0.95 * p.min_tag_width * p.min_tag_width) bad = true;
        }
        if (!bad) {  // :209-238
          for (int i = 0; i < 4; i++) {
            const int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
            const float dx1 = cr[i1][0] - cr[i0][0], dy1 = cr[i1][1] - cr[i0][1];
            const float dx2 = cr[i2][0] - cr[i1][0], dy2 = cr[i2][1] - cr[i1][1];
            const float cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrtf((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
            if (fabsf(cos_dtheta) > p.cos_critical_rad || dx1 * dy2 < dy1 * dx2) { bad = true; break; }
          }
        }
        if (!bad) {
          const float f = static_cast<float>(p.f);
          if (f > 1) {  // AdjustPixelCenters, :260-282
            for (int j = 0; j < 4; j++) {
              cr[j][0] = (cr[j][0] - 0.5f) * f + 0.5f;
              cr[j][1] = (cr[j][1] - 0.5f) * f + 0.5f;
            }
          }
          const uint32_t qi = atomicAdd(&ctr->num_quads, 1u);
          if (qi < p.quad_cap) {
            b200tag_quad &q = quads[qi];
            for (int j = 0; j < 4; j++) { q.corners[j][0] = cr[j][0]; q.corners[j][1] = cr[j][1]; }
            q.reversed_border = p.reversed_border && !p.normal_border;
            q.blob_index = b;
            q.rep0 = blob.rep0;
            q.rep1 = blob.rep1;
          } else {
            atomicOr(&ctr->status, B200TAG_ST_QUADS_OVERFLOW);
          }
        }
      }
    }
  }
}

int launch_blobs(const FrameParams &p, int frames, cudaStream_t s, KernelTimer *kt) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_fit_blobs, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FitShared)));
    attr_set = true;
  }
  if (kt) kt->begin("select", s);
  k_select<<<dim3(min(p.hash_cap / 256u, 296u), frames), 256, 0, s>>>(p);
  if (kt) kt->end(s);
  if (kt) kt->begin("scatter", s);
  k_scatter<<<dim3(592, frames), 256, 0, s>>>(p);
  if (kt) kt->end(s);
  if (kt) kt->begin("fit_blobs", s);
  k_fit_blobs<<<dim3(296, frames), kFitThreads, sizeof(FitShared), s>>>(p);
  if (kt) kt->end(s);
  return 3;
}

}  // namespace b200tag
