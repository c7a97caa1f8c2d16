// Baseline-JPEG luminance decoder on the GPU (camera wire format).  The reference leaves MJPG
// decoding to OpenCV on the CPU (src/usb_camera/src/camera_publisher.cpp:198,336) and then converts bgr8 -> YUYV ->
// gray; the detector only ever looks at luminance, so this kernel decodes exactly that plane, straight into the
// detector's input staging buffer.
//
// Huffman-coded data is sequential within a restart interval (ITU-T T.81 F.2.2), so the parallelism is across frames:
// one warp per frame.  All 32 lanes run the entropy decoder redundantly on identical state (same instructions, same
// data: no divergence, no cost over a single thread), which keeps every branch warp-uniform and lets the lanes split
// the data-parallel parts: byte unstuffing into a shared-memory ring (ballot / popc compaction), the 8x8 inverse DCT
// (two coefficients per lane) and the pixel stores.  Chrominance blocks are parsed and dropped.
#include <cuda_runtime.h>

#include <cstdint>

#include "jpeg.h"
#include "kernels.h"

namespace b200tag {
namespace {

constexpr int kRingBytes = 1024;  // unstuffed entropy-coded bytes staged in shared memory (power of two)
constexpr int kRingWords = kRingBytes / 4;

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Shared {
  JpegTables tab;
  JpegFrame frame;
  uint32_t ring[kRingWords];
  float coef[64];  // dequantised coefficients of the current block, natural order
  float tmp[64];
  float cosv[64];  // cos((2x+1) u pi / 16) * C(u) / 2, [x][u]
  float quant[64];
  uint8_t zigzag[64];
};

// Input side: raw bytes -> ring of unstuffed bytes (T.81 B.1.1.5: FF 00 stands for a data byte FF; FF followed by
// anything else is a marker and ends the segment).
struct Feeder {
  const uint8_t *src;
  uint32_t gpos, end;   // next raw byte, end of the JPEG
  uint32_t wr;          // bytes written to the ring (monotonic)
  uint32_t prev;        // last raw byte of the previous window (for an FF / 00 pair split across windows)
  bool stopped;         // a marker (or the end of the data) was reached at gpos
};

__device__ __forceinline__ void feed(Feeder &f, uint32_t rd_words, uint8_t *ring_bytes, int lane) {
  // keep the ring topped up: at least kRingBytes - 64 bytes ahead of the reader while input lasts; zero padding after
  // the end of the segment (a well-formed interval never reads it)
  while (f.wr - 4u * rd_words <= static_cast<uint32_t>(kRingBytes - 64)) {
    if (f.stopped) {
      ring_bytes[(f.wr + lane) & (kRingBytes - 1)] = 0;
      f.wr += 32;
      continue;
    }
    const uint32_t idx = f.gpos + lane;
    const uint32_t b = idx < f.end ? f.src[idx] : 0xffu;               // past the end reads as a marker (FF D9)
    uint32_t nb = __shfl_down_sync(0xffffffffu, b, 1);
    if (lane == 31) nb = idx + 1 < f.end ? f.src[idx + 1] : 0xd9u;
    uint32_t pb = __shfl_up_sync(0xffffffffu, b, 1);
    if (lane == 0) pb = f.prev;
    const bool marker = (b == 0xffu && nb != 0u) || idx >= f.end;
    const bool drop = (pb == 0xffu && b == 0u);
    const uint32_t mmask = __ballot_sync(0xffffffffu, marker);
    const int first = mmask ? __ffs(mmask) - 1 : 32;
    const bool keep = lane < first && !drop;
    const uint32_t kmask = __ballot_sync(0xffffffffu, keep);
    if (keep) ring_bytes[(f.wr + __popc(kmask & ((1u << lane) - 1u))) & (kRingBytes - 1)] = static_cast<uint8_t>(b);
    f.wr += __popc(kmask);
    if (first < 32) {
      f.stopped = true;
      f.gpos += first;
    } else {
      f.gpos += 32;
      f.prev = __shfl_sync(0xffffffffu, b, 31);
    }
  }
  __syncwarp();
}

struct BitReader {
  unsigned long long acc;
  int nbits;
  uint32_t rd;  // ring words consumed
};

__device__ __forceinline__ void ensure(BitReader &r, const uint32_t *ring) {
  if (r.nbits <= 32) {
    const uint32_t w = ring[r.rd & (kRingWords - 1)];
    r.acc = (r.acc << 32) | __byte_perm(w, 0, 0x0123);
    r.nbits += 32;
    r.rd++;
  }
}

// F.2.2.3 DECODE: 9-bit lookahead, canonical bounds for longer codes.  Needs nbits >= 16.
__device__ __forceinline__ uint32_t decode_symbol(BitReader &r, const JpegHuff &h) {
  const uint32_t peek = static_cast<uint32_t>(r.acc >> (r.nbits - 16)) & 0xffffu;
  const uint32_t e = h.fast[peek >> (16 - kJpegFastBits)];
  if (e) {
    r.nbits -= static_cast<int>(e >> 8);
    return e & 0xffu;
  }
  int l = kJpegFastBits + 1;
  while (l <= 16 && static_cast<int>(peek >> (16 - l)) > h.maxcode[l]) l++;
  if (l > 16) {  // not a code of this table (corrupt stream)
    r.nbits -= 16;
    return 0;
  }
  r.nbits -= l;
  return h.vals[(h.valoff[l] + static_cast<int>(peek >> (16 - l))) & 0xff];
}

// F.2.2.1 RECEIVE + EXTEND; s <= 11 bits, available after decode_symbol without another refill
__device__ __forceinline__ int receive_extend(BitReader &r, uint32_t s) {
  if (s == 0) return 0;
  const int v = static_cast<int>(static_cast<uint32_t>(r.acc >> (r.nbits - static_cast<int>(s))) & ((1u << s) - 1u));
  r.nbits -= static_cast<int>(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

__global__ void __launch_bounds__(32) k_jpeg_luma(const uint8_t *__restrict__ bits, const JpegFrame *__restrict__ frames,
                                                  const JpegTables *__restrict__ tables, uint8_t *__restrict__ out,
                                                  size_t out_stride) {
  __shared__ Shared S;
  const int lane = threadIdx.x;
  static_assert(sizeof(JpegFrame) % 4 == 0 && sizeof(JpegTables) % 4 == 0, "copied as words");
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames + blockIdx.x);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&S.frame);
    for (uint32_t i = lane; i < sizeof(JpegFrame) / 4; i += 32) dst[i] = src[i];
  }
  __syncwarp();
  const JpegFrame *F = &S.frame;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(tables + F->tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&S.tab);
    for (uint32_t i = lane; i < sizeof(JpegTables) / 4; i += 32) dst[i] = src[i];
    for (int i = lane; i < 64; i += 32) {
      S.quant[i] = static_cast<float>(F->quant[i]);
      S.zigzag[i] = c_zigzag[i];
      const int x = i >> 3, u = i & 7;
      S.cosv[i] = cospif(static_cast<float>((2 * x + 1) * u) * 0.0625f) * (u == 0 ? 0.70710678118654752f : 1.0f) * 0.5f;
    }
  }
  const int width = F->width, height = F->height;
  const int mcus_x = F->mcus_x, nmcu = F->mcus_x * F->mcus_y;
  const int nblocks = F->nblocks, hmax = F->hmax, vmax = F->vmax;
  const int restart = F->restart_interval;
  uint8_t *img = out + static_cast<size_t>(blockIdx.x) * out_stride;
  uint8_t *ring_bytes = reinterpret_cast<uint8_t *>(S.ring);

  Feeder fd;
  fd.src = bits + F->data_off;
  fd.gpos = 0;
  fd.end = F->data_len;
  fd.wr = 0;
  fd.prev = 0x100u;
  fd.stopped = false;
  BitReader br;
  br.acc = 0;
  br.nbits = 0;
  br.rd = 0;
  int pred0 = 0, pred1 = 0, pred2 = 0;  // DC predictors (F.2.1.3.1)
  int until_restart = restart;
  __syncwarp();

  int mx = 0, my = 0;
  for (int m = 0; m < nmcu; m++) {
    if (restart && until_restart == 0) {
      // E.2.4: the interval is followed by an RSTn marker; find it (the feeder normally stopped exactly there), step
      // over it and start the next interval on a byte boundary with zero predictors
      while (!fd.stopped) {
        fd.wr = 4u * br.rd;  // discard
        feed(fd, br.rd, ring_bytes, lane);
        br.rd = fd.wr / 4u;  // (feed only returns with the ring nearly full or the marker found)
      }
      const uint32_t mk = fd.gpos + 1 < fd.end ? fd.src[fd.gpos + 1] : 0xd9u;
      if (fd.gpos < fd.end && mk >= 0xd0u && mk <= 0xd7u) {
        fd.gpos += 2;
        fd.stopped = false;
        fd.prev = 0x100u;
      }
      fd.wr = 0;
      br.rd = 0;
      br.acc = 0;
      br.nbits = 0;
      pred0 = pred1 = pred2 = 0;
      until_restart = restart;
      __syncwarp();
    }
    until_restart--;
    for (int blk = 0; blk < nblocks; blk++) {
      feed(fd, br.rd, ring_bytes, lane);
      const int comp = F->blk_comp[blk];
      const bool luma = comp == 0;
      const JpegHuff &hdc = S.tab.dc[F->comp_dc[comp]];
      const JpegHuff &hac = S.tab.ac[F->comp_ac[comp]];
      if (luma) {
        S.coef[lane] = 0.0f;
        S.coef[lane + 32] = 0.0f;
      }
      __syncwarp();
      ensure(br, S.ring);
      const uint32_t t = decode_symbol(br, hdc) & 15u;
      ensure(br, S.ring);
      const int diff = receive_extend(br, t > 11u ? 11u : t);
      if (comp == 0) pred0 += diff;
      else if (comp == 1) pred1 += diff;
      else pred2 += diff;
      bool any_ac = false;
      if (luma && lane == 0) S.coef[0] = static_cast<float>(pred0) * S.quant[0];
      for (int k = 1; k < 64;) {
        ensure(br, S.ring);
        const uint32_t rs = decode_symbol(br, hac);
        const uint32_t run = rs >> 4, s = rs & 15u;
        if (s == 0) {
          if (run != 15u) break;  // EOB
          k += 16;                // ZRL
          continue;
        }
        k += static_cast<int>(run);
        if (k > 63) break;
        const int v = receive_extend(br, s > 10u ? 10u : s);
        if (luma && lane == 0) {
          const int nat = S.zigzag[k];
          S.coef[nat] = static_cast<float>(v) * S.quant[nat];
        }
        any_ac = true;
        k++;
      }
      if (!luma) continue;
      __syncwarp();
      const int bx0 = (mx * hmax + F->blk_bx[blk]) * 8, by0 = (my * vmax + F->blk_by[blk]) * 8;
      if (!any_ac) {  // DC only: a flat block
        const float v = S.coef[0] * 0.125f + 128.0f;
        const int pv = min(255, max(0, __float2int_rn(v)));
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int o = lane + 32 * h, px = bx0 + (o & 7), py = by0 + (o >> 3);
          if (px < width && py < height) img[static_cast<size_t>(py) * width + px] = static_cast<uint8_t>(pv);
        }
        continue;
      }
      // A.3.3 inverse DCT, separable: columns (tmp[y][u] = sum_v c[y][v] coef[v][u]) then rows
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int o = lane + 32 * h, y = o >> 3, u = o & 7;
        float s = 0.0f;
#pragma unroll
        for (int v = 0; v < 8; v++) s += S.cosv[y * 8 + v] * S.coef[v * 8 + u];
        S.tmp[o] = s;
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int o = lane + 32 * h, y = o >> 3, x = o & 7;
        float s = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u++) s += S.cosv[x * 8 + u] * S.tmp[y * 8 + u];
        const int pv = min(255, max(0, __float2int_rn(s + 128.0f)));
        const int px = bx0 + x, py = by0 + y;
        if (px < width && py < height) img[static_cast<size_t>(py) * width + px] = static_cast<uint8_t>(pv);
      }
    }
    if (++mx == mcus_x) {
      mx = 0;
      my++;
    }
  }
}

}  // namespace

void launch_jpeg_luma(const uint8_t *bits, const JpegFrame *frames, const JpegTables *tables, uint8_t *out, size_t out_stride,
                      int count, cudaStream_t s) {
  k_jpeg_luma<<<count, 32, 0, s>>>(bits, frames, tables, out, out_stride);
}

}  // namespace b200tag
