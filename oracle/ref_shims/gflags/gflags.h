// TEST INFRASTRUCTURE: stand-in for gflags (flags become plain globals with their defaults).
#ifndef ORC_SHIM_GFLAGS_H_
#define ORC_SHIM_GFLAGS_H_
#include <cstdint>
#define DEFINE_bool(name, def, help) bool FLAGS_##name = def
#define DEFINE_int32(name, def, help) int32_t FLAGS_##name = def
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#endif
