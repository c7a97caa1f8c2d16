#ifndef ORC_SHIM_G2D_H_
#define ORC_SHIM_G2D_H_
#include "apriltag.h"
#ifdef __cplusplus
extern "C" {
#endif
zarray_t *g2d_polygon_create_zeros(int sz);
#ifdef __cplusplus
}
#endif
#endif
