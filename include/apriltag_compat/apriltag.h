/*
 * Stand-in for the libapriltag headers (apriltag.h, common/zarray.h, common/matd.h,
 * common/image_u8.h, common/workerpool.h, tag36h11.h) that the reference's callers include
 * (src/apriltags_cuda/include/apriltags_cuda/apriltag_gpu.h:6,
 *  src/apriltags_cuda/src/apriltags_cuda_detector.cu:139-147,425-433,
 *  src/apriltags_cuda/test/gpu_detector_test.cu:11-13).
 *
 * libapriltag (github.com/cgpadwick/apriltag tag 3.3.0, src/external/CMakeLists.txt:86-95) is not
 * vendored in the reference tree and not installed here, so these declarations let the
 * GpuDetector class in include/apriltags_cuda/apriltag_gpu.h compile and be tested stand-alone.
 * Struct layouts follow upstream AprilTag 3.x (RECALLED).  In the real
 * workspace put libapriltag's include directory first on the include path and these are not used.
 */
#ifndef B200TAG_APRILTAG_COMPAT_H_
#define B200TAG_APRILTAG_COMPAT_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- common/zarray.h ---- */
typedef struct zarray zarray_t;
struct zarray {
  size_t el_sz; /* size of each element */
  int size;     /* how many elements? */
  int alloc;    /* we've allocated storage for how many elements? */
  char *data;
};

static inline zarray_t *zarray_create(size_t el_sz) {
  zarray_t *za = (zarray_t *)calloc(1, sizeof(zarray_t));
  za->el_sz = el_sz;
  return za;
}
static inline void zarray_destroy(zarray_t *za) {
  if (za == NULL) return;
  if (za->data != NULL) free(za->data);
  memset(za, 0, sizeof(zarray_t));
  free(za);
}
static inline int zarray_size(const zarray_t *za) { return za->size; }
static inline void zarray_ensure_capacity(zarray_t *za, int capacity) {
  if (capacity <= za->alloc) return;
  while (za->alloc < capacity) {
    za->alloc *= 2;
    if (za->alloc < 8) za->alloc = 8;
  }
  za->data = (char *)realloc(za->data, za->alloc * za->el_sz);
}
static inline void zarray_add(zarray_t *za, const void *p) {
  zarray_ensure_capacity(za, za->size + 1);
  memcpy(&za->data[za->size * za->el_sz], p, za->el_sz);
  za->size++;
}
static inline void zarray_get(const zarray_t *za, int idx, void *p) { memcpy(p, &za->data[idx * za->el_sz], za->el_sz); }
static inline void zarray_set(zarray_t *za, int idx, const void *p, void *outp) {
  if (outp != NULL) memcpy(outp, &za->data[idx * za->el_sz], za->el_sz);
  memcpy(&za->data[idx * za->el_sz], p, za->el_sz);
}
static inline void zarray_truncate(zarray_t *za, int sz) { za->size = sz; }
static inline void zarray_clear(zarray_t *za) { za->size = 0; }
static inline void zarray_sort(zarray_t *za, int (*compar)(const void *, const void *)) {
  if (za->size == 0) return;
  qsort(za->data, za->size, za->el_sz, compar);
}

/* ---- common/matd.h ---- */
typedef struct {
  unsigned int nrows, ncols;
  double data[];
} matd_t;
#define MATD_EL(m, row, col) (m)->data[((row) * (m)->ncols + (col))]
matd_t *matd_create(int rows, int cols);
matd_t *matd_create_data(int rows, int cols, const double *data);
void matd_destroy(matd_t *m);
static inline double matd_get(const matd_t *m, unsigned int row, unsigned int col) { return MATD_EL(m, row, col); }

/* ---- common/image_u8.h ---- */
typedef struct image_u8 image_u8_t;
struct image_u8 {
  const int32_t width;
  const int32_t height;
  const int32_t stride;
  uint8_t *buf;
};

/* ---- common/workerpool.h (opaque) ---- */
typedef struct workerpool workerpool_t;
workerpool_t *workerpool_create(int nthreads);
void workerpool_destroy(workerpool_t *wp);

/* ---- apriltag.h ---- */
typedef struct apriltag_family apriltag_family_t;
struct apriltag_family {
  uint32_t ncodes;
  uint64_t *codes;
  int width_at_border;
  int total_width;
  bool reversed_border;
  uint32_t nbits;
  uint32_t *bit_x;
  uint32_t *bit_y;
  uint32_t h;
  char *name;
  void *impl;
};

struct apriltag_quad_thresh_params {
  int min_cluster_pixels;
  int max_nmaxima;
  float critical_rad;
  float cos_critical_rad;
  float max_line_fit_mse;
  int min_white_black_diff;
  int deglitch;
};

typedef struct apriltag_detector apriltag_detector_t;
struct apriltag_detector {
  int nthreads;
  float quad_decimate;
  float quad_sigma;
  bool refine_edges;
  double decode_sharpening;
  bool debug;
  struct apriltag_quad_thresh_params qtp;
  void *tp; /* timeprofile_t* */
  uint32_t nedges;
  uint32_t nsegments;
  uint32_t nquads;
  zarray_t *tag_families;
  workerpool_t *wp;
  void *mutex_storage[8]; /* pthread_mutex_t in libapriltag */
};

typedef struct apriltag_detection apriltag_detection_t;
struct apriltag_detection {
  apriltag_family_t *family;
  int id;
  int hamming;
  float decision_margin;
  matd_t *H;
  double c[2];
  double p[4][2];
};

apriltag_detector_t *apriltag_detector_create(void);
void apriltag_detector_add_family_bits(apriltag_detector_t *td, apriltag_family_t *fam, int bits_corrected);
static inline void apriltag_detector_add_family(apriltag_detector_t *td, apriltag_family_t *fam) {
  apriltag_detector_add_family_bits(td, fam, 2);
}
void apriltag_detector_destroy(apriltag_detector_t *td);
void apriltag_detection_destroy(apriltag_detection_t *det);
void apriltag_detections_destroy(zarray_t *detections);

/* ---- apriltag_pose.h ---- */
typedef struct {
  apriltag_detection_t *det;
  double tagsize; /* in meters */
  double fx, fy, cx, cy;
} apriltag_detection_info_t;
typedef struct {
  matd_t *R; /* 3x3 */
  matd_t *t; /* 3x1 */
} apriltag_pose_t;
/* The call the node makes per detection (apriltags_cuda_detector.cu:433).  Implemented in
 * libapriltags_cuda_core.so over b200tag_estimate_pose, so the node's pose step no longer needs libapriltag. */
double estimate_tag_pose(apriltag_detection_info_t *info, apriltag_pose_t *pose);

/* ---- tag36h11.h ---- */
apriltag_family_t *tag36h11_create(void);
void tag36h11_destroy(apriltag_family_t *tf);
/* ---- tag25h9.h, tag16h5.h ---- */
apriltag_family_t *tag25h9_create(void);
void tag25h9_destroy(apriltag_family_t *tf);
apriltag_family_t *tag16h5_create(void);
void tag16h5_destroy(apriltag_family_t *tf);

#ifdef __cplusplus
}
#endif
#endif
