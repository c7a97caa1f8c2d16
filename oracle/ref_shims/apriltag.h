// TEST INFRASTRUCTURE: libapriltag declarations the reference's apriltag_gpu.cu / apriltag_detect.cu need,
// layered on the repo's stand-in header.  The decode entry points are stubbed in oracle/ref_harness.cu:
// oracle/_ref exercises the reference GPU front end through QuadCorners only.
#ifndef ORC_SHIM_APRILTAG_H_
#define ORC_SHIM_APRILTAG_H_
#include "../../include/apriltag_compat/apriltag.h"
#ifdef __cplusplus
extern "C" {
#endif
#define APRILTAG_TASKS_PER_THREAD_TARGET 10
struct quad {
  float p[4][2];
  bool reversed_border;
  matd_t *H, *Hinv;
};
void workerpool_add_task(workerpool_t *wp, void (*f)(void *p), void *p);
void workerpool_run(workerpool_t *wp);
image_u8_t *image_u8_copy(const image_u8_t *in);
void image_u8_darken(image_u8_t *im);
void image_u8_draw_line(image_u8_t *im, float x0, float y0, float x1, float y1, int v, int width);
int image_u8_write_pnm(const image_u8_t *im, const char *path);
#ifdef __cplusplus
}
#endif
#endif
