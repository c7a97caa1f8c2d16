// Baseline-JPEG luminance decoder (camera wire format): structures shared by the host-side
// header parser (jpeg_host.cc) and the decode kernel (kernels_jpeg.cu).  ITU-T T.81 section / figure numbers in the
// comments refer to the JPEG standard; the reference itself leaves MJPG decoding to OpenCV on the CPU
// (src/usb_camera/src/camera_publisher.cpp:198,336).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200tag {

constexpr int kJpegFastBits = 9;
constexpr int kJpegMaxBlocksPerMcu = 10;  // T.81 B.2.3

// One Huffman table in decoder form (T.81 F.2.2.3): a 9-bit lookahead table for the short codes and the canonical
// per-length bounds for the rest.
struct JpegHuff {
  uint16_t fast[1 << kJpegFastBits];  // (code length << 8) | symbol; 0 = code longer than 9 bits
  int32_t maxcode[18];                // largest code of each length 1..16 (-1: none); [17] is a sentinel
  int32_t valoff[18];                 // valptr[l] - mincode[l]
  uint8_t vals[256];
};

struct JpegTables {
  JpegHuff dc[2], ac[2];
};

// One frame of a batch, as the kernel needs it.
struct JpegFrame {
  uint32_t data_off;  // entropy-coded segment: offset into the batch's bitstream buffer ...
  uint32_t data_len;  // ... and bytes up to the end of the JPEG
  uint16_t width, height;
  uint16_t mcus_x, mcus_y;
  uint16_t restart_interval;  // MCUs between RSTn markers, 0 = none
  uint8_t nblocks;            // blocks per MCU
  uint8_t tables;             // index of the JpegTables set of this frame
  uint8_t hmax, vmax;         // luminance blocks per MCU, horizontally / vertically
  uint8_t blk_comp[kJpegMaxBlocksPerMcu];  // component (0 = luminance) of each block of an MCU, in stream order
  uint8_t blk_bx[kJpegMaxBlocksPerMcu];    // position of a luminance block inside its MCU, in blocks
  uint8_t blk_by[kJpegMaxBlocksPerMcu];
  uint8_t comp_dc[4], comp_ac[4];          // Huffman table selectors per component
  uint16_t quant[64];                      // luminance quantisation table, natural (row-major) order
  uint32_t chunk_off;                      // first unstuffing chunk / first subsequence of this frame in the batch arrays
  uint32_t sub_off;
};

// Device workspace of one batch for the parallel decoder (kernels_jpeg.cu)
struct JpegSyncState;
struct JpegBatch {
  const uint8_t *raw;         // descriptors, tables and entropy-coded segments as uploaded
  uint8_t *clean;             // unstuffed segments, at the same offsets as in `raw`
  const JpegFrame *frames;
  const JpegTables *tables;
  uint32_t *chunk_cnt;        // per 64-byte raw chunk: bytes kept, then their exclusive prefix
  uint32_t *clean_len;        // per frame: unstuffed bytes
  uint32_t *chunk_rst;        // per chunk: RSTn markers ending in it, then their exclusive prefix
  uint32_t *nrst;             // per frame: RSTn markers found
  uint32_t *rst_pos;          // [frame][rst_stride]: unstuffed byte offset at which the interval after each marker starts
  uint32_t rst_stride;
  uint32_t max_intervals;     // largest number of restart intervals of a frame in the batch
  unsigned long long *sync;   // per subsequence: JpegSyncState at its start
  unsigned long long *sync_in; // per subsequence: the start state its published successor state was decoded from
  uint32_t *nblk;             // per subsequence: blocks completed in it, then their exclusive prefix
  uint32_t *changed;          // [frame][round]: some state changed in that round
  uint32_t *proven;           // per frame: the parallel decode reached its fixed point
  int16_t *coef;              // [frame][luminance block][64], zigzag order; all zero between batches
  size_t coef_stride;         // int16 per frame
  int16_t *dcs;               // [frame][luminance block]: DC differences, then DC values (coef_stride / 64 per frame)
  uint8_t *out;
  size_t out_stride;
  int count;
  uint32_t max_chunks, max_subs, max_luma_blocks;  // largest per-frame counts in the batch (grid sizes)
};

// Host side ---------------------------------------------------------------------------------------------------------
enum { kJpegOk = 0, kJpegMalformed = -1, kJpegUnsupported = -2 };

struct JpegParsed {
  JpegFrame frame;              // data_off / tables still to be filled by the caller
  size_t scan_begin = 0;        // first byte of the entropy-coded segment within the file
  std::vector<uint8_t> dht;     // canonical form of the four tables in use (counts + values), to find identical sets
};

// Parses the headers of one JPEG (markers up to SOS).  kJpegUnsupported: a valid JPEG this decoder does not handle
// (progressive, arithmetic, 12-bit, non-interleaved scans, subsampled luminance, table ids above 1).
int jpeg_parse(const uint8_t *data, size_t len, JpegParsed *out, std::string *why);
// Decoder form of the tables in JpegParsed::dht (once per distinct set of a batch); false: not a valid prefix code
bool jpeg_build_tables(const std::vector<uint8_t> &dht, JpegTables *out);

// Host model of the parallel decoder (same core, same arithmetic, threads replaced by loops) -- a test hook that lets
// the scheme be checked without a GPU; nothing on the detection path calls it.  Returns 0, 1 = stream kind the
// parallel path does not take (restart markers, non-baseline), negative = malformed.  rounds = synchronisation rounds
// until the proving round (Jacobi order, an upper bound for the kernels' in-place order).
int jpeg_model_decode(const uint8_t *jpeg, size_t len, uint8_t *out, size_t out_cap, int *rounds);

}  // namespace b200tag
