/* Minimal implementations of the libapriltag entry points declared in
 * include/apriltag_compat/apriltag.h, so the GpuDetector class can be built and tested without
 * libapriltag.  Only object construction / destruction: detection itself runs in libb200tag.so.
 * Defaults follow upstream apriltag_detector_create() (AprilTag 3.x, recalled). */
#include <math.h>
#include <stdio.h>

#include "apriltag.h"
#include "tag_families_data.h"

matd_t *matd_create(int rows, int cols) {
  matd_t *m = (matd_t *)calloc(1, sizeof(matd_t) + (size_t)rows * cols * sizeof(double));
  m->nrows = (unsigned)rows;
  m->ncols = (unsigned)cols;
  return m;
}
matd_t *matd_create_data(int rows, int cols, const double *data) {
  matd_t *m = matd_create(rows, cols);
  memcpy(m->data, data, (size_t)rows * cols * sizeof(double));
  return m;
}
void matd_destroy(matd_t *m) { free(m); }

struct workerpool { int nthreads; };
workerpool_t *workerpool_create(int nthreads) {
  workerpool_t *wp = (workerpool_t *)calloc(1, sizeof(workerpool_t));
  wp->nthreads = nthreads;
  return wp;
}
void workerpool_destroy(workerpool_t *wp) { free(wp); }

apriltag_detector_t *apriltag_detector_create(void) {
  apriltag_detector_t *td = (apriltag_detector_t *)calloc(1, sizeof(apriltag_detector_t));
  td->nthreads = 1;
  td->quad_decimate = 2.0f;
  td->quad_sigma = 0.0f;
  td->qtp.max_nmaxima = 10;
  td->qtp.min_cluster_pixels = 5;
  td->qtp.max_line_fit_mse = 10.0f;
  td->qtp.cos_critical_rad = cosf((float)(10 * M_PI / 180));
  td->qtp.critical_rad = (float)(10 * M_PI / 180);
  td->qtp.deglitch = 0;
  td->qtp.min_white_black_diff = 5;
  td->tag_families = zarray_create(sizeof(apriltag_family_t *));
  td->refine_edges = 1;
  td->decode_sharpening = 0.25;
  td->debug = 0;
  td->wp = workerpool_create(1);
  return td;
}
void apriltag_detector_add_family_bits(apriltag_detector_t *td, apriltag_family_t *fam, int bits_corrected) {
  (void)bits_corrected;
  zarray_add(td->tag_families, &fam);
}
void apriltag_detector_destroy(apriltag_detector_t *td) {
  if (!td) return;
  workerpool_destroy(td->wp);
  zarray_destroy(td->tag_families);
  free(td);
}
void apriltag_detection_destroy(apriltag_detection_t *det) {
  if (det == NULL) return;
  matd_destroy(det->H);
  free(det);
}
void apriltag_detections_destroy(zarray_t *detections) {
  for (int i = 0; i < zarray_size(detections); i++) {
    apriltag_detection_t *det;
    zarray_get(detections, i, &det);
    apriltag_detection_destroy(det);
  }
  zarray_destroy(detections);
}

static apriltag_family_t *family_create(const char *name, uint32_t h, uint32_t nbits, uint32_t ncodes, const uint64_t *codes,
                                        const int32_t *bit_x, const int32_t *bit_y, int width_at_border, int total_width) {
  apriltag_family_t *tf = (apriltag_family_t *)calloc(1, sizeof(apriltag_family_t));
  tf->name = strdup(name);
  tf->h = h;
  tf->ncodes = ncodes;
  tf->codes = (uint64_t *)calloc(ncodes, sizeof(uint64_t));
  memcpy(tf->codes, codes, ncodes * sizeof(uint64_t));
  tf->nbits = nbits;
  tf->bit_x = (uint32_t *)calloc(nbits, sizeof(uint32_t));
  tf->bit_y = (uint32_t *)calloc(nbits, sizeof(uint32_t));
  for (uint32_t i = 0; i < nbits; i++) {
    tf->bit_x[i] = (uint32_t)bit_x[i];
    tf->bit_y[i] = (uint32_t)bit_y[i];
  }
  tf->width_at_border = width_at_border;
  tf->total_width = total_width;
  tf->reversed_border = false;
  return tf;
}
static void family_destroy(apriltag_family_t *tf) {
  if (!tf) return;
  free(tf->codes);
  free(tf->bit_x);
  free(tf->bit_y);
  free(tf->name);
  free(tf);
}

#define B200_FAMILY(NAME)                                                                                              \
  apriltag_family_t *NAME##_create(void) {                                                                             \
    return family_create(#NAME, b200_##NAME##_MIN_HAMMING, b200_##NAME##_NBITS, b200_##NAME##_NCODES, b200_##NAME##_codes, \
                         b200_##NAME##_bit_x, b200_##NAME##_bit_y, b200_##NAME##_WIDTH_AT_BORDER, b200_##NAME##_TOTAL_WIDTH); \
  }                                                                                                                    \
  void NAME##_destroy(apriltag_family_t *tf) { family_destroy(tf); }

B200_FAMILY(tag36h11)
B200_FAMILY(tag25h9)
B200_FAMILY(tag16h5)
