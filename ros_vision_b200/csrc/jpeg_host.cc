// Host side of the JPEG luminance decoder: marker parsing (ITU-T T.81 B.2), Huffman decoder tables (Annex C, F.2.2.3).
// A few hundred header bytes per frame; the entropy-coded data is not touched here.
#include "jpeg.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "jpeg_core.h"

namespace b200tag {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// T.81 tables K.3 - K.6 ("typical" Huffman tables), implied by MJPG streams that carry no DHT segment.  The AC value
// lists are generated: within each code length the standard orders the run/size symbols as listed here.
struct RawTable {
  uint8_t counts[16];
  std::vector<uint8_t> vals;
  bool present = false;
};

const uint8_t kDcLumCounts[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrCounts[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kAcLumCounts[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125};
const uint8_t kAcChrCounts[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119};
// first 32 / 43 symbols (codes of up to 11 / 14 bits) are irregular; the long tail is every remaining run/size pair in
// increasing order
const uint8_t kAcLumHead[] = {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
                              0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
                              0x24, 0x33, 0x62, 0x72, 0x82};
const uint8_t kAcChrHead[] = {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
                              0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
                              0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1};

RawTable standard_ac(const uint8_t *counts, const uint8_t *head, size_t nhead) {
  RawTable t;
  memcpy(t.counts, counts, 16);
  bool used[256] = {false};
  for (size_t i = 0; i < nhead; i++) {
    t.vals.push_back(head[i]);
    used[head[i]] = true;
  }
  for (int rs = 0; rs < 256; rs++) {  // the remaining run/size symbols: sizes 1..10, plus nothing else
    const int s = rs & 15;
    if (used[rs] || s < 1 || s > 10) continue;
    t.vals.push_back(static_cast<uint8_t>(rs));
  }
  t.present = true;
  return t;
}

RawTable standard_dc(const uint8_t *counts) {
  RawTable t;
  memcpy(t.counts, counts, 16);
  for (int i = 0; i < 12; i++) t.vals.push_back(static_cast<uint8_t>(i));
  t.present = true;
  return t;
}

bool build_huff(const RawTable &raw, JpegHuff *h) {
  memset(h, 0, sizeof(*h));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    const int n = raw.counts[l - 1];
    h->valoff[l] = k - code;
    if (n) {
      if (code + n > (1 << l)) return false;  // over-subscribed code
      for (int i = 0; i < n; i++) {
        if (l <= kJpegFastBits) {
          const int first = (code + i) << (kJpegFastBits - l), span = 1 << (kJpegFastBits - l);
          for (int j = 0; j < span; j++) h->fast[first + j] = static_cast<uint16_t>((l << 8) | raw.vals[k + i]);
        }
      }
      h->maxcode[l] = code + n - 1;
    } else {
      h->maxcode[l] = -1;
    }
    code = (code + n) << 1;
    k += n;
  }
  h->maxcode[17] = 0x7fffffff;
  if (k > 256) return false;
  memcpy(h->vals, raw.vals.data(), raw.vals.size());
  return true;
}

int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

}  // namespace

int jpeg_parse(const uint8_t *data, size_t len, JpegParsed *out, std::string *why) {
  auto fail = [&](int rc, const char *msg) {
    if (why) *why = msg;
    return rc;
  };
  if (!data || len < 4 || data[0] != 0xff || data[1] != 0xd8) return fail(kJpegMalformed, "no SOI marker");
  RawTable dc[4], ac[4];
  uint16_t quant[4][64];
  bool have_q[4] = {false, false, false, false};
  int width = 0, height = 0, ncomp = 0, hs[4] = {0}, vs[4] = {0}, tq[4] = {0}, cid[4] = {0};
  int restart = 0;
  bool have_sof = false;
  size_t pos = 2;
  int td[4] = {0}, ta[4] = {0};
  for (;;) {
    if (pos + 4 > len || data[pos] != 0xff) return fail(kJpegMalformed, "marker expected");
    while (pos < len && data[pos] == 0xff) pos++;
    if (pos >= len) return fail(kJpegMalformed, "truncated");
    const int m = data[pos++];
    if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
    if (m == 0xd9) return fail(kJpegMalformed, "EOI before SOS");
    if (pos + 2 > len) return fail(kJpegMalformed, "truncated");
    const int L = be16(data + pos);
    if (L < 2 || pos + static_cast<size_t>(L) > len) return fail(kJpegMalformed, "bad segment length");
    const uint8_t *seg = data + pos + 2;
    const int n = L - 2;
    if (m == 0xc0 || m == 0xc1) {
      if (n < 6) return fail(kJpegMalformed, "short SOF");
      if (seg[0] != 8) return fail(kJpegUnsupported, "sample precision is not 8 bits");
      height = be16(seg + 1);
      width = be16(seg + 3);
      ncomp = seg[5];
      if (ncomp != 1 && ncomp != 3) return fail(kJpegUnsupported, "neither 1 nor 3 components");
      if (n < 6 + 3 * ncomp) return fail(kJpegMalformed, "short SOF");
      for (int c = 0; c < ncomp; c++) {
        cid[c] = seg[6 + 3 * c];
        hs[c] = seg[7 + 3 * c] >> 4;
        vs[c] = seg[7 + 3 * c] & 15;
        tq[c] = seg[8 + 3 * c] & 3;
        if (hs[c] < 1 || hs[c] > 4 || vs[c] < 1 || vs[c] > 4) return fail(kJpegMalformed, "bad sampling factor");
      }
      have_sof = true;
    } else if (m >= 0xc2 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
      return fail(kJpegUnsupported, "not a baseline (sequential, Huffman) JPEG");
    } else if (m == 0xc4) {
      int o = 0;
      while (o + 17 <= n) {
        const int tc = seg[o] >> 4, th = seg[o] & 15;
        if (tc > 1 || th > 3) return fail(kJpegMalformed, "bad DHT");
        int total = 0;
        for (int i = 0; i < 16; i++) total += seg[o + 1 + i];
        if (total > 256 || o + 17 + total > n) return fail(kJpegMalformed, "bad DHT");
        RawTable &t = tc ? ac[th] : dc[th];
        memcpy(t.counts, seg + o + 1, 16);
        t.vals.assign(seg + o + 17, seg + o + 17 + total);
        t.present = true;
        o += 17 + total;
      }
    } else if (m == 0xdb) {
      int o = 0;
      while (o < n) {
        const int pq = seg[o] >> 4, t = seg[o] & 15;
        if (t > 3 || pq > 1 || o + 1 + (pq ? 128 : 64) > n) return fail(kJpegMalformed, "bad DQT");
        for (int i = 0; i < 64; i++) quant[t][kZigzag[i]] = pq ? static_cast<uint16_t>(be16(seg + o + 1 + 2 * i)) : seg[o + 1 + i];
        have_q[t] = true;
        o += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xdd) {
      if (n < 2) return fail(kJpegMalformed, "bad DRI");
      restart = be16(seg);
    } else if (m == 0xda) {
      if (!have_sof || n < 1) return fail(kJpegMalformed, "SOS before SOF");
      const int ns = seg[0];
      if (n < 1 + 2 * ns + 3) return fail(kJpegMalformed, "short SOS");
      if (ns != ncomp) return fail(kJpegUnsupported, "non-interleaved scans");
      for (int i = 0; i < ns; i++) {
        if (seg[1 + 2 * i] != cid[i]) return fail(kJpegUnsupported, "scan components out of order");
        td[i] = seg[2 + 2 * i] >> 4;
        ta[i] = seg[2 + 2 * i] & 15;
        if (td[i] > 1 || ta[i] > 1) return fail(kJpegUnsupported, "Huffman table ids above 1");
      }
      if (seg[1 + 2 * ns] != 0 || seg[2 + 2 * ns] != 63) return fail(kJpegUnsupported, "spectral selection (progressive scan)");
      pos += static_cast<size_t>(L);
      break;
    }
    pos += static_cast<size_t>(L);
  }
  if (width < 1 || height < 1) return fail(kJpegMalformed, "empty frame");
  if (!have_q[tq[0]]) return fail(kJpegMalformed, "luminance quantisation table missing");
  bool any = false;
  for (int i = 0; i < 2; i++) any = any || dc[i].present || ac[i].present;
  if (!any) {  // no DHT at all: tables K.3 - K.6
    dc[0] = standard_dc(kDcLumCounts);
    dc[1] = standard_dc(kDcChrCounts);
    ac[0] = standard_ac(kAcLumCounts, kAcLumHead, sizeof(kAcLumHead));
    ac[1] = standard_ac(kAcChrCounts, kAcChrHead, sizeof(kAcChrHead));
  }
  int hmax = 1, vmax = 1;
  if (ncomp == 1) {
    hs[0] = vs[0] = 1;  // T.81 A.2.2: a single-component scan is never interleaved
  } else {
    for (int c = 0; c < ncomp; c++) {
      hmax = hs[c] > hmax ? hs[c] : hmax;
      vmax = vs[c] > vmax ? vs[c] : vmax;
    }
    if (hs[0] != hmax || vs[0] != vmax) return fail(kJpegUnsupported, "subsampled luminance");
  }
  JpegFrame &f = out->frame;
  memset(&f, 0, sizeof(f));
  int nb = 0;
  for (int c = 0; c < ncomp; c++) {
    if (!dc[td[c]].present || !ac[ta[c]].present) return fail(kJpegMalformed, "scan refers to a missing Huffman table");
    for (int by = 0; by < vs[c]; by++)
      for (int bx = 0; bx < hs[c]; bx++) {
        if (nb >= kJpegMaxBlocksPerMcu) return fail(kJpegMalformed, "more than 10 blocks per MCU");
        f.blk_comp[nb] = static_cast<uint8_t>(c);
        f.blk_bx[nb] = static_cast<uint8_t>(bx);
        f.blk_by[nb] = static_cast<uint8_t>(by);
        nb++;
      }
    f.comp_dc[c] = static_cast<uint8_t>(td[c]);
    f.comp_ac[c] = static_cast<uint8_t>(ta[c]);
  }
  f.nblocks = static_cast<uint8_t>(nb);
  f.width = static_cast<uint16_t>(width);
  f.height = static_cast<uint16_t>(height);
  f.hmax = static_cast<uint8_t>(hmax);
  f.vmax = static_cast<uint8_t>(vmax);
  f.mcus_x = static_cast<uint16_t>((width + 8 * hmax - 1) / (8 * hmax));
  f.mcus_y = static_cast<uint16_t>((height + 8 * vmax - 1) / (8 * vmax));
  if (restart > 0xffff) return fail(kJpegMalformed, "bad DRI");
  f.restart_interval = static_cast<uint16_t>(restart);
  memcpy(f.quant, quant[tq[0]], sizeof(f.quant));
  for (int i = 0; i < 64; i++)
    if (f.quant[i] == 0) return fail(kJpegMalformed, "zero quantiser");
  out->scan_begin = pos;
  out->dht.clear();
  for (int i = 0; i < 2; i++)
    for (const RawTable *t : {&dc[i], &ac[i]}) {
      out->dht.push_back(t->present ? 1 : 0);
      if (!t->present) continue;
      out->dht.insert(out->dht.end(), t->counts, t->counts + 16);
      out->dht.insert(out->dht.end(), t->vals.begin(), t->vals.end());
    }
  return kJpegOk;
}

bool jpeg_build_tables(const std::vector<uint8_t> &dht, JpegTables *out) {
  memset(out, 0, sizeof(*out));
  size_t o = 0;
  for (int i = 0; i < 2; i++)
    for (int kind = 0; kind < 2; kind++) {  // order of jpeg_parse: dc[i], ac[i]
      if (o >= dht.size()) return false;
      if (!dht[o++]) continue;
      if (o + 16 > dht.size()) return false;
      RawTable t;
      memcpy(t.counts, dht.data() + o, 16);
      size_t total = 0;
      for (int l = 0; l < 16; l++) total += t.counts[l];
      o += 16;
      if (total > 256 || o + total > dht.size()) return false;
      t.vals.assign(dht.begin() + o, dht.begin() + o + total);
      o += total;
      if (!build_huff(t, kind ? &out->ac[i] : &out->dc[i])) return false;
    }
  return true;
}

int jpeg_model_decode(const uint8_t *jpeg, size_t len, uint8_t *out, size_t out_cap, int *rounds) {
  JpegParsed P;
  std::string why;
  const int rc = jpeg_parse(jpeg, len, &P, &why);
  if (rc == kJpegUnsupported) return 1;
  if (rc != kJpegOk) return -1;
  const JpegFrame &F = P.frame;
  JpegTables tables;
  if (!jpeg_build_tables(P.dht, &tables)) return -1;
  if (out_cap < static_cast<size_t>(F.width) * F.height) return -1;
  // unstuffing passes (k_jpeg_unstuff_*): FF 00 -> FF, RSTn markers out (their positions = interval starts)
  std::vector<uint8_t> clean;
  std::vector<uint32_t> rst;
  for (size_t i = P.scan_begin; i < len; i++) {
    const uint8_t b = jpeg[i], prev = i > P.scan_begin ? jpeg[i - 1] : 0, nb = i + 1 < len ? jpeg[i + 1] : 0;
    if (b == 0 && prev == 0xff) continue;
    if (b == 0xff && nb >= 0xd0 && nb <= 0xd7) continue;
    if (prev == 0xff && b >= 0xd0 && b <= 0xd7) {
      rst.push_back(static_cast<uint32_t>(clean.size()));
      continue;
    }
    clean.push_back(b);
  }
  const uint32_t end_bits = static_cast<uint32_t>(clean.size()) * 8u;
  clean.resize(((clean.size() + 3) & ~size_t(3)) + 16, 0);
  const uint32_t *words = reinterpret_cast<const uint32_t *>(clean.data());
  const uint32_t luma_per_mcu = static_cast<uint32_t>(F.hmax) * F.vmax, nmcu = static_cast<uint32_t>(F.mcus_x) * F.mcus_y;
  std::vector<int16_t> coef(static_cast<size_t>(nmcu) * luma_per_mcu * 64, 0), dcs(static_cast<size_t>(nmcu) * luma_per_mcu, 0);
  const uint32_t nsub = F.restart_interval ? 0 : (end_bits + kJpegSubBits - 1) / kJpegSubBits;
  if (F.restart_interval) {  // k_jpeg_write_rst: one "thread" per restart interval
    const uint32_t ri = F.restart_interval, nint = (nmcu + ri - 1) / ri;
    if (rst.size() + 1 != nint) return -2;
    for (uint32_t k = 0; k < nint; k++) {
      JpegSyncState st{k ? rst[k - 1] * 8u : 0u, 0u};
      JpegIntervalSink sink{{coef.data(), dcs.data(), k * ri, nmcu, luma_per_mcu, F.nblocks}, 0};
      jpeg_decode_span(words, end_bits, k + 1 < nint ? rst[k] * 8u : end_bits, F, tables, st, sink, std::min(ri, nmcu - k * ri) * F.nblocks);
    }
    if (rounds) *rounds = 0;
  }
  // synchronisation rounds (k_jpeg_sync)
  std::vector<JpegSyncState> s(nsub ? nsub : 1, JpegSyncState{0, 0}), next;
  std::vector<uint32_t> nblk(nsub ? nsub : 1, 0);
  JpegNullSink none;
  int used = 0;
  bool proven = nsub <= 1;
  for (int r = 0; r < 4096 && !proven && !F.restart_interval; r++) {
    next = s;
    bool changed = false;
    for (uint32_t i = 0; i + 1 < nsub; i++) {
      JpegSyncState st = (r == 0 || i == 0) ? JpegSyncState{i * kJpegSubBits, 0} : s[i];
      nblk[i] = jpeg_decode_span(words, end_bits, (i + 1) * kJpegSubBits, F, tables, st, none);
      if (r == 0 || st.pos != s[i + 1].pos || st.cz != s[i + 1].cz) {
        next[i + 1] = st;
        changed = true;
      }
    }
    s.swap(next);
    used = r + 1;
    if (!changed) proven = true;
  }
  if (rounds && !F.restart_interval) *rounds = used;
  if (!proven) return -2;
  // block numbering (k_jpeg_blockscan) and coefficient pass (k_jpeg_write)
  uint32_t base = 0;
  for (uint32_t i = 0; i < nsub; i++) {
    JpegSyncState st = i == 0 ? JpegSyncState{0, 0} : s[i];
    JpegCoefSink sink{coef.data(), dcs.data(), base / F.nblocks, nmcu, luma_per_mcu, F.nblocks};
    const uint32_t done = jpeg_decode_span(words, end_bits, i + 1 == nsub ? 0xffffffffu : (i + 1) * kJpegSubBits, F, tables, st, sink);
    base += i + 1 == nsub ? done : nblk[i];
  }
  // DC prediction (k_jpeg_dcscan) and inverse DCT (k_jpeg_idct)
  int pred = 0;
  for (uint32_t lb = 0; lb < nmcu * luma_per_mcu; lb++) {
    int16_t *zz = &coef[static_cast<size_t>(lb) * 64];
    pred = F.restart_interval ? dcs[lb] : pred + dcs[lb];
    float nat[64], tmp[64];
    bool any_ac = false;
    for (int k = 0; k < 64; k++) {
      const int v = k == 0 ? pred : zz[k];
      any_ac = any_ac || (k > 0 && v != 0);
      nat[kZigzag[k]] = static_cast<float>(v) * static_cast<float>(F.quant[kZigzag[k]]);
    }
    const uint32_t mcu = lb / luma_per_mcu, j = lb % luma_per_mcu;
    const int bx0 = static_cast<int>((mcu % F.mcus_x) * F.hmax + F.blk_bx[j]) * 8;
    const int by0 = static_cast<int>((mcu / F.mcus_x) * F.vmax + F.blk_by[j]) * 8;
    for (int u = 0; u < 8; u++) {  // columns, then rows (k_jpeg_idct)
      float in[8], o[8];
      for (int v = 0; v < 8; v++) in[v] = nat[v * 8 + u];
      jpeg_idct8(in, o);
      for (int y = 0; y < 8; y++) tmp[y * 8 + u] = o[y];
    }
    for (int y = 0; y < 8; y++) {
      float o[8];
      jpeg_idct8(tmp + y * 8, o);
      for (int x = 0; x < 8; x++) {
        float v = any_ac ? o[x] + 128.0f : nat[0] * 0.125f + 128.0f;
        const int pv = static_cast<int>(std::nearbyintf(v));
        const int px = bx0 + x, py = by0 + y;
        if (px < F.width && py < F.height) out[static_cast<size_t>(py) * F.width + px] = static_cast<uint8_t>(pv < 0 ? 0 : (pv > 255 ? 255 : pv));
      }
    }
  }
  return 0;
}

}  // namespace b200tag
