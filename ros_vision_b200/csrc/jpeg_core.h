// Entropy-decoding core of the JPEG luminance decoder, shared by the CUDA kernels (kernels_jpeg.cu) and by the host
// model of the parallel scheme (jpeg_host.cc, a test hook that mirrors the kernels thread by thread).
//
// Parallel Huffman decoding by self-synchronisation (Klein & Wiseman 2003; Weissenberger & Schmidt 2018 for JPEG on
// GPUs): the unstuffed scan is cut into subsequences of kJpegSubBits bits.  The decoder state at a subsequence
// boundary is (bit position of the first codeword that starts in the subsequence, block index inside the MCU, zigzag
// index inside the block).  f_i maps the state at the start of subsequence i to the state at the start of i + 1 by
// decoding.  Thread i iterates  s[i+1] <- f_i(s[i])  starting from the guess "a block starts at my first bit"; because
// Huffman decoders that start in different states fall into step after a few codewords, the iteration reaches its
// fixed point -- which, s[0] being known, is the true decoding -- after a few rounds instead of after n.  A round in
// which nothing changes proves the fixed point.
#pragma once
#include <cstdint>

#include "jpeg.h"

#if defined(__CUDACC__)
#define JPEG_HD __host__ __device__ __forceinline__
#else
#define JPEG_HD inline
#endif

namespace b200tag {

constexpr uint32_t kJpegSubBits = 512;   // bits per subsequence (64 bytes of unstuffed data)
constexpr int kJpegSyncRounds = 8;       // rounds launched; a frame that is not proven by then is decoded sequentially
constexpr uint32_t kJpegChunk = 64;      // raw bytes per thread of the unstuffing passes

struct JpegSyncState {
  uint32_t pos;  // bit offset of the next codeword in the unstuffed scan
  uint32_t cz;   // block index within the MCU | zigzag index << 8 (0 = a DC codeword is next)
};

// 32 bits of the big-endian bit stream starting at bit `pos`; `words` is 4-byte aligned
JPEG_HD uint32_t jpeg_peek32(const uint32_t *words, uint32_t pos) {
  const uint32_t i = pos >> 5, sh = pos & 31u;
  uint32_t w0 = words[i], w1 = words[i + 1];
#if defined(__CUDA_ARCH__)
  w0 = __byte_perm(w0, 0, 0x0123);
  w1 = __byte_perm(w1, 0, 0x0123);
  return __funnelshift_l(w1, w0, sh);
#else
  w0 = __builtin_bswap32(w0);
  w1 = __builtin_bswap32(w1);
  return sh ? (w0 << sh) | (w1 >> (32u - sh)) : w0;
#endif
}

// F.2.2.3 DECODE on the 16 leading bits of `peek`; returns the symbol, *len = code length
JPEG_HD uint32_t jpeg_lookup(const JpegHuff &h, uint32_t peek, uint32_t *len) {
  const uint32_t p16 = peek >> 16;
  const uint32_t e = h.fast[p16 >> (16 - kJpegFastBits)];
  if (e) {
    *len = e >> 8;
    return e & 0xffu;
  }
  uint32_t l = kJpegFastBits + 1;
  while (l <= 16 && static_cast<int32_t>(p16 >> (16 - l)) > h.maxcode[l]) l++;
  if (l > 16) {  // not a codeword of this table
    *len = 16;
    return 0;
  }
  *len = l;
  return h.vals[(h.valoff[l] + static_cast<int32_t>(p16 >> (16 - l))) & 0xff];
}

// F.2.2.1 RECEIVE + EXTEND: the s bits that follow the `len`-bit codeword in `peek`
JPEG_HD int jpeg_value(uint32_t peek, uint32_t len, uint32_t s) {
  if (s == 0) return 0;
  const int v = static_cast<int>((peek << len) >> (32u - s));
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// Decodes codewords from state `s` until one ends at or beyond bit `limit`, the data ends at `end_bits`, or
// `max_blocks` blocks are complete; returns the number of blocks completed.  sink.dc(c, diff), sink.ac(c, z, v) and sink.block_end(c) see every coefficient.
JPEG_HD uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// big-endian word `i` of the bit stream
JPEG_HD uint32_t jpeg_word(const uint32_t *words, uint32_t i) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(words[i], 0, 0x0123);
#else
  return __builtin_bswap32(words[i]);
#endif
}

template <class Sink>
JPEG_HD uint32_t jpeg_decode_span(const uint32_t *words, uint32_t end_bits, uint32_t limit, const JpegFrame &F,
                                  const JpegTables &T, JpegSyncState &s, Sink &sink, uint32_t max_blocks = 0xffffffffu) {
  uint32_t pos = s.pos, c = s.cz & 0xffu, z = s.cz >> 8, blocks = 0;
  const uint32_t nblocks = F.nblocks;
  // bit buffer: the next `nbits` (> 32 before every codeword) bits of the stream, left aligned; a codeword and its
  // value bits (<= 27) are taken from the top half, one word is fetched per 32 bits consumed
  uint32_t wi = (pos >> 5) + 2;
  unsigned long long acc = ((static_cast<unsigned long long>(jpeg_word(words, wi - 2)) << 32) | jpeg_word(words, wi - 1)) << (pos & 31u);
  int nbits = 64 - static_cast<int>(pos & 31u);
  // Huffman tables of the current block's component (they change at block ends only)
  const JpegHuff *hdc = &T.dc[F.comp_dc[F.blk_comp[c]]], *hac = &T.ac[F.comp_ac[F.blk_comp[c]]];
  while (pos < limit && pos < end_bits && blocks < max_blocks) {
    if (nbits <= 32) {
      acc |= static_cast<unsigned long long>(jpeg_word(words, wi++)) << (32 - nbits);
      nbits += 32;
    }
    const uint32_t peek = static_cast<uint32_t>(acc >> 32);
    // one code path for DC and AC codewords (the lanes of a warp are at different places of their blocks): F.2.2.1 /
    // F.2.2.2 differ in the table, in the run and in what the decoded value is
    const bool is_dc = z == 0;
    uint32_t len;
    const uint32_t rs = jpeg_lookup(is_dc ? *hdc : *hac, peek, &len);
    const uint32_t run = is_dc ? 0u : rs >> 4;
    uint32_t sz = rs & 15u;
    sz = min_u32(sz, is_dc ? 11u : 10u);
    const int v = jpeg_value(peek, len, sz);
    const uint32_t used = len + sz;
    if (is_dc) {
      sink.dc(c, v);
      z = 1;
    } else if (sz == 0) {
      z = run == 15u ? z + 16u : 64u;  // ZRL : EOB
    } else {
      z += run;
      if (z <= 63u) sink.ac(c, z, v);
      z++;
    }
    pos += used;
    acc <<= used;
    nbits -= static_cast<int>(used);
    if (z >= 64u) {
      sink.block_end(c);
      z = 0;
      blocks++;
      if (++c == nblocks) c = 0;
      const uint32_t comp = F.blk_comp[c];
      hdc = &T.dc[F.comp_dc[comp]];
      hac = &T.ac[F.comp_ac[comp]];
    }
  }
  s.pos = pos;
  s.cz = c | (z << 8);
  return blocks;
}

// 8-point inverse DCT of T.81 A.3.3, x[n] = sum_k C(k)/2 X[k] cos((2n+1) k pi / 16), by the even/odd split (22
// multiplications and 28 additions instead of 64 + 56).  One function for the kernels and the host model: the same
// operations in the same order, no contraction into FMAs (the library is built with -fmad=false / -ffp-contract=off).
JPEG_HD void jpeg_idct8(const float X[8], float x[8]) {
  constexpr float d1 = 0.49039264f, d2 = 0.46193977f, d3 = 0.41573481f, d4 = 0.35355339f, d5 = 0.27778512f, d6 = 0.19134172f,
                  d7 = 0.09754516f;  // cos(j pi / 16) / 2
  const float ee0 = (X[0] + X[4]) * d4, ee1 = (X[0] - X[4]) * d4;
  const float eo0 = X[2] * d2 + X[6] * d6, eo1 = X[2] * d6 - X[6] * d2;
  const float e0 = ee0 + eo0, e1 = ee1 + eo1, e2 = ee1 - eo1, e3 = ee0 - eo0;
  const float o0 = X[1] * d1 + X[3] * d3 + X[5] * d5 + X[7] * d7;
  const float o1 = X[1] * d3 - X[3] * d7 - X[5] * d1 - X[7] * d5;
  const float o2 = X[1] * d5 - X[3] * d1 + X[5] * d7 + X[7] * d3;
  const float o3 = X[1] * d7 - X[3] * d5 + X[5] * d3 - X[7] * d1;
  x[0] = e0 + o0; x[7] = e0 - o0;
  x[1] = e1 + o1; x[6] = e1 - o1;
  x[2] = e2 + o2; x[5] = e2 - o2;
  x[3] = e3 + o3; x[4] = e3 - o3;
}

struct JpegNullSink {
  JPEG_HD void dc(uint32_t, int) {}
  JPEG_HD void ac(uint32_t, uint32_t, int) {}
  JPEG_HD void block_end(uint32_t) {}
};

// Writes the luminance coefficients (zigzag order, DC as the difference to the previous luminance block) of the blocks
// it sees, numbered from `block` (count of all blocks of all components before the span).
struct JpegCoefSink {
  int16_t *coef;       // [luminance block][64]; entry 0 of a block is not used ...
  int16_t *dcs;        // ... the DC values have an array of their own, [luminance block]
  uint32_t mcu;        // MCU of the block being decoded
  uint32_t nmcu;
  uint32_t luma_per_mcu;
  uint32_t nblocks;
  JPEG_HD void dc(uint32_t c, int v) {
    if (c < luma_per_mcu && mcu < nmcu) dcs[static_cast<size_t>(mcu) * luma_per_mcu + c] = static_cast<int16_t>(v);
  }
  JPEG_HD void ac(uint32_t c, uint32_t z, int v) {
    if (c < luma_per_mcu && mcu < nmcu) coef[(static_cast<size_t>(mcu) * luma_per_mcu + c) * 64 + z] = static_cast<int16_t>(v);
  }
  JPEG_HD void block_end(uint32_t c) {
    if (c + 1 == nblocks) mcu++;
  }
};

// The same for one restart interval decoded from its start by a single thread: DC prediction (F.2.1.3.1) is a running
// sum inside the interval, so the absolute value is written straight away.
struct JpegIntervalSink {
  JpegCoefSink to;
  int pred;
  JPEG_HD void dc(uint32_t c, int v) {
    if (c < to.luma_per_mcu) {
      pred += v;
      to.dc(c, pred);
    }
  }
  JPEG_HD void ac(uint32_t c, uint32_t z, int v) { to.ac(c, z, v); }
  JPEG_HD void block_end(uint32_t c) { to.block_end(c); }
};

}  // namespace b200tag
