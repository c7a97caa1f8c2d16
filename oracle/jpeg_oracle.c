/* TEST INFRASTRUCTURE -- CPU restatement of baseline JPEG decoding, luminance plane only.
 *
 * The reference has no JPEG code of its own: its camera node asks OpenCV (cv::VideoCapture with CAP_PROP_CONVERT_RGB,
 * src/usb_camera/src/camera_publisher.cpp:198,336) to decode the cameras' MJPG stream, i.e. the libjpeg-turbo bundled
 * with OpenCV, which the reference pins at 4.9.0 (src/external/CMakeLists.txt:27-35, fetched at build time) -- a
 * third-party dependency that is not vendored in the reference tree.  This file restates the published
 * algorithm, ITU-T T.81 (baseline sequential DCT, Huffman coding): marker parsing (B.2), Huffman table generation
 * (Annex C), decoding of DC / AC coefficients (F.2.2), dequantisation and the inverse DCT (A.3.3), level shift (A.3.1).
 * JPEG decoders are only required to agree within the accuracy bounds of T.83, not bit for bit; the oracle is pinned
 * against the libjpeg-turbo in this image's OpenCV (4.13.0 / libjpeg-turbo 3.1.2; tests/test_jpeg_oracle.py: within
 * 1 grey level on every pixel of every test stream).  Straightforward on purpose: bit-by-bit canonical Huffman decoding, the IDCT as the
 * double-precision double sum of A.3.3.  Only tests/, smoke() and bench.py's CPU legs may use anything in oracle/. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int bits[17];       /* number of codes of each length 1..16 */
  uint8_t vals[256];
  int mincode[17], maxcode[18], valptr[17];
  int present;
} huff_t;

typedef struct {
  const uint8_t *p, *end;
  uint32_t acc;
  int n;
  int marker; /* marker byte that stopped the entropy-coded segment, 0 if none yet */
} bits_t;

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* Tables K.3 - K.6: what a stream without DHT segments (plain UVC / AVI MJPG) implies */
static const uint8_t kStdDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kStdDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kStdDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kStdAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kStdAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kStdAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kStdAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

/* Annex C: code sizes -> codes -> decoder tables (F.2.2.3, figure F.15) */
static void huff_build(huff_t *h) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    h->valptr[l] = k;
    h->mincode[l] = code;
    code += h->bits[l];
    k += h->bits[l];
    h->maxcode[l] = h->bits[l] ? code - 1 : -1;
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  h->present = 1;
}

static void huff_set(huff_t *h, const uint8_t *bits16, const uint8_t *vals, int nvals) {
  memset(h, 0, sizeof(*h));
  for (int i = 0; i < 16; i++) h->bits[i + 1] = bits16[i];
  memcpy(h->vals, vals, (size_t)nvals);
  huff_build(h);
}

/* F.2.2.5 NEXTBIT with byte unstuffing (B.1.1.5); past the end of the segment the stream reads as zero bits */
static int next_bit(bits_t *b) {
  if (b->n == 0) {
    uint32_t v = 0;
    if (!b->marker && b->p < b->end) {
      v = *b->p++;
      if (v == 0xff) {
        const int m = b->p < b->end ? *b->p : 0xd9;
        if (m == 0) b->p++;
        else { b->marker = m; b->p--; v = 0; }
      }
    }
    b->acc = v;
    b->n = 8;
  }
  b->n--;
  return (int)((b->acc >> b->n) & 1u);
}

static int receive(bits_t *b, int s) {
  int v = 0;
  for (int i = 0; i < s; i++) v = (v << 1) | next_bit(b);
  return v;
}

static int extend(int v, int t) { return t == 0 ? 0 : (v < (1 << (t - 1)) ? v - (1 << t) + 1 : v); } /* F.12 */

static int decode_symbol(bits_t *b, const huff_t *h) { /* F.16 */
  int code = next_bit(b), l = 1;
  while (l <= 16 && code > h->maxcode[l]) {
    code = (code << 1) | next_bit(b);
    l++;
  }
  if (l > 16) return 0; /* corrupt stream: treat as symbol 0 */
  return h->vals[h->valptr[l] + code - h->mincode[l]];
}

typedef struct {
  int width, height, ncomp, hs[4], vs[4], tq[4], cid[4], restart_interval;
} jpeg_info_t;

static int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

/* Decodes the luminance (first) component of a baseline JPEG into out[height][width]; returns 0, or a negative code:
 * -1 malformed, -2 not baseline 8-bit Huffman, -3 output too small.  `info` (optional) receives the frame header. */
int jpeg_oracle_decode_luma(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, jpeg_info_t *info) {
  huff_t dc[4], ac[4];
  uint16_t quant[4][64];
  int have_q[4] = {0, 0, 0, 0};
  jpeg_info_t fi;
  memset(&fi, 0, sizeof(fi));
  memset(dc, 0, sizeof(dc));
  memset(ac, 0, sizeof(ac));
  if (len < 4 || data[0] != 0xff || data[1] != 0xd8) return -1;
  size_t pos = 2;
  int sos_td[4] = {0}, sos_ta[4] = {0}, sos_n = 0, sos_comp[4] = {0};
  int have_sof = 0;
  for (;;) {
    if (pos + 4 > len) return -1;
    if (data[pos] != 0xff) return -1;
    while (pos < len && data[pos] == 0xff) pos++; /* fill bytes */
    if (pos >= len) return -1;
    const int m = data[pos++];
    if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
    if (m == 0xd9) return -1;
    if (pos + 2 > len) return -1;
    const int L = be16(data + pos);
    if (L < 2 || pos + (size_t)L > len) return -1;
    const uint8_t *seg = data + pos + 2;
    const int n = L - 2;
    if (m == 0xc0 || m == 0xc1) { /* SOF0 / SOF1 with 8-bit samples */
      if (n < 6 || seg[0] != 8) return -2;
      fi.height = be16(seg + 1);
      fi.width = be16(seg + 3);
      fi.ncomp = seg[5];
      if (fi.ncomp < 1 || fi.ncomp > 4 || n < 6 + 3 * fi.ncomp) return -1;
      for (int c = 0; c < fi.ncomp; c++) {
        fi.cid[c] = seg[6 + 3 * c];
        fi.hs[c] = seg[7 + 3 * c] >> 4;
        fi.vs[c] = seg[7 + 3 * c] & 15;
        fi.tq[c] = seg[8 + 3 * c] & 3;
        if (fi.hs[c] < 1 || fi.hs[c] > 4 || fi.vs[c] < 1 || fi.vs[c] > 4) return -1;
      }
      have_sof = 1;
    } else if (m >= 0xc2 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
      return -2; /* progressive, lossless, arithmetic ... */
    } else if (m == 0xc4) { /* DHT */
      int o = 0;
      while (o + 17 <= n) {
        const int tc = seg[o] >> 4, th = seg[o] & 15;
        if (tc > 1 || th > 3) return -1;
        int total = 0;
        for (int i = 0; i < 16; i++) total += seg[o + 1 + i];
        if (total > 256 || o + 17 + total > n) return -1;
        huff_set(tc ? &ac[th] : &dc[th], seg + o + 1, seg + o + 17, total);
        o += 17 + total;
      }
    } else if (m == 0xdb) { /* DQT */
      int o = 0;
      while (o < n) {
        const int pq = seg[o] >> 4, tq = seg[o] & 15;
        if (tq > 3 || o + 1 + (pq ? 128 : 64) > n) return -1;
        for (int i = 0; i < 64; i++) quant[tq][kZigzag[i]] = pq ? (uint16_t)be16(seg + o + 1 + 2 * i) : seg[o + 1 + i];
        have_q[tq] = 1;
        o += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xdd) { /* DRI */
      if (n < 2) return -1;
      fi.restart_interval = be16(seg);
    } else if (m == 0xda) { /* SOS */
      if (!have_sof || n < 1) return -1;
      sos_n = seg[0];
      if (sos_n != fi.ncomp || n < 1 + 2 * sos_n + 3) return -2; /* one interleaved scan only */
      for (int i = 0; i < sos_n; i++) {
        sos_comp[i] = seg[1 + 2 * i];
        sos_td[i] = seg[2 + 2 * i] >> 4;
        sos_ta[i] = seg[2 + 2 * i] & 15;
        if (sos_comp[i] != fi.cid[i] || sos_td[i] > 3 || sos_ta[i] > 3) return -2;
      }
      if (seg[1 + 2 * sos_n] != 0 || seg[2 + 2 * sos_n] != 63) return -2;
      pos += (size_t)L;
      break;
    }
    pos += (size_t)L;
  }
  if (info) *info = fi;
  if (fi.width < 1 || fi.height < 1) return -1;
  if (!out) return 0;
  if (out_cap < (size_t)fi.width * fi.height) return -3;
  if (!have_q[fi.tq[0]]) return -1;
  if (!dc[0].present && !dc[1].present && !ac[0].present && !ac[1].present) { /* no DHT at all: tables K.3 - K.6 */
    huff_set(&dc[0], kStdDcLumBits, kStdDcVals, 12);
    huff_set(&dc[1], kStdDcChrBits, kStdDcVals, 12);
    huff_set(&ac[0], kStdAcLumBits, kStdAcLumVals, 162);
    huff_set(&ac[1], kStdAcChrBits, kStdAcChrVals, 162);
  }
  for (int i = 0; i < sos_n; i++)
    if (!dc[sos_td[i]].present || !ac[sos_ta[i]].present) return -1;

  int hmax = 1, vmax = 1;
  for (int c = 0; c < fi.ncomp; c++) {
    if (fi.hs[c] > hmax) hmax = fi.hs[c];
    if (fi.vs[c] > vmax) vmax = fi.vs[c];
  }
  if (fi.ncomp == 1) { fi.hs[0] = fi.vs[0] = 1; hmax = vmax = 1; } /* A.2.2: a single component is never interleaved */
  if (fi.hs[0] != hmax || fi.vs[0] != vmax) return -2; /* luminance at full resolution */
  const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
  const int mcus_x = (fi.width + mcu_w - 1) / mcu_w, mcus_y = (fi.height + mcu_h - 1) / mcu_h;

  double cosv[8][8]; /* cos((2x+1) u pi / 16) * C(u) / 2 */
  for (int x = 0; x < 8; x++)
    for (int u = 0; u < 8; u++) cosv[x][u] = cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0) * (u == 0 ? sqrt(0.5) : 1.0) * 0.5;

  bits_t b = {data + pos, data + len, 0, 0, 0};
  int pred[4] = {0, 0, 0, 0};
  int until_restart = fi.restart_interval;
  for (int my = 0; my < mcus_y; my++) {
    for (int mx = 0; mx < mcus_x; mx++) {
      if (fi.restart_interval && until_restart == 0) { /* F.2.2.4 / E.2.4: the RSTn marker between intervals */
        b.n = 0;
        if (!b.marker) { /* the decoder stopped short of the marker: skip to it */
          while (b.p + 1 < b.end && !(b.p[0] == 0xff && b.p[1] >= 0xd0 && b.p[1] <= 0xd7)) b.p++;
        }
        if (b.p + 1 < b.end && b.p[0] == 0xff && b.p[1] >= 0xd0 && b.p[1] <= 0xd7) b.p += 2;
        b.marker = 0;
        pred[0] = pred[1] = pred[2] = pred[3] = 0;
        until_restart = fi.restart_interval;
      }
      until_restart--;
      for (int c = 0; c < fi.ncomp; c++) {
        for (int by = 0; by < fi.vs[c]; by++) {
          for (int bx = 0; bx < fi.hs[c]; bx++) {
            int coef[64];
            memset(coef, 0, sizeof(coef));
            const int t = decode_symbol(&b, &dc[sos_td[c]]);
            pred[c] += extend(receive(&b, t), t);
            coef[0] = pred[c];
            for (int k = 1; k < 64;) { /* F.2.2.2 */
              const int rs = decode_symbol(&b, &ac[sos_ta[c]]);
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break; /* EOB */
                k += 16;            /* ZRL */
                continue;
              }
              k += r;
              if (k > 63) break;
              coef[kZigzag[k]] = extend(receive(&b, s), s);
              k++;
            }
            if (c != 0) continue; /* chrominance is parsed and dropped */
            double deq[64], tmp[64];
            for (int i = 0; i < 64; i++) deq[i] = (double)coef[i] * quant[fi.tq[0]][i];
            for (int y = 0; y < 8; y++)     /* A.3.3, rows then columns */
              for (int u = 0; u < 8; u++) {
                double s = 0;
                for (int v = 0; v < 8; v++) s += cosv[y][v] * deq[v * 8 + u];
                tmp[y * 8 + u] = s;
              }
            for (int y = 0; y < 8; y++)
              for (int x = 0; x < 8; x++) {
                double s = 0;
                for (int u = 0; u < 8; u++) s += cosv[x][u] * tmp[y * 8 + u];
                const int px = (mx * hmax + bx) * 8 + x, py = (my * vmax + by) * 8 + y;
                if (px < fi.width && py < fi.height) {
                  double v = floor(s + 128.0 + 0.5);
                  out[(size_t)py * fi.width + px] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
                }
              }
          }
        }
      }
    }
  }
  return 0;
}

/* the K.3 - K.6 tables as one DHT payload (Tc/Th, 16 counts, values ...), for tests that compare them with libjpeg's */
int jpeg_oracle_standard_dht(uint8_t *out, size_t cap) {
  const uint8_t *bits[4] = {kStdDcLumBits, kStdAcLumBits, kStdDcChrBits, kStdAcChrBits};
  const uint8_t *vals[4] = {kStdDcVals, kStdAcLumVals, kStdDcVals, kStdAcChrVals};
  const int nv[4] = {12, 162, 12, 162}, id[4] = {0x00, 0x10, 0x01, 0x11};
  size_t o = 0;
  for (int t = 0; t < 4; t++) {
    if (o + 17 + (size_t)nv[t] > cap) return -1;
    out[o++] = (uint8_t)id[t];
    memcpy(out + o, bits[t], 16); o += 16;
    memcpy(out + o, vals[t], (size_t)nv[t]); o += (size_t)nv[t];
  }
  return (int)o;
}
