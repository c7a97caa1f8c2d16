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
};

// Host side ---------------------------------------------------------------------------------------------------------
enum { kJpegOk = 0, kJpegMalformed = -1, kJpegUnsupported = -2 };

struct JpegParsed {
  JpegFrame frame;              // data_off / tables still to be filled by the caller
  size_t scan_begin = 0;        // first byte of the entropy-coded segment within the file
  std::vector<uint8_t> dht;     // canonical form of the four tables in use (counts + values), to find identical sets
  JpegTables tables;
};

// Parses the headers of one JPEG (markers up to SOS).  kJpegUnsupported: a valid JPEG this decoder does not handle
// (progressive, arithmetic, 12-bit, non-interleaved scans, subsampled luminance, table ids above 1).
int jpeg_parse(const uint8_t *data, size_t len, JpegParsed *out, std::string *why);

}  // namespace b200tag
