// TEST INFRASTRUCTURE: minimal stand-in for glog so the reference's CUDA sources compile unmodified
// from /root/reference (oracle/build_ref.sh).  LOG(FATAL) / failed CHECKs abort; everything else is swallowed.
#ifndef ORC_SHIM_GLOG_LOGGING_H_
#define ORC_SHIM_GLOG_LOGGING_H_
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string_view>
#include "gflags/gflags.h"  // real glog pulls gflags in; apriltag_detect.cu:20 relies on it
namespace orc_shim {
struct NullStream {
  template <typename T> NullStream &operator<<(const T &) { return *this; }
  NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
struct FatalStream {
  std::ostringstream ss;
  template <typename T> FatalStream &operator<<(const T &v) { ss << v; return *this; }
  FatalStream &operator<<(std::ostream &(*f)(std::ostream &)) { ss << f; return *this; }
  [[noreturn]] ~FatalStream() {
    std::fprintf(stderr, "reference FATAL: %s\n", ss.str().c_str());
    std::abort();
  }
};
struct Voidify {
  void operator&(NullStream &) {}
  void operator&(FatalStream &) {}
};
constexpr int INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3;
template <int S> struct Pick { using type = NullStream; };
template <> struct Pick<FATAL> { using type = FatalStream; };
}  // namespace orc_shim
#define LOG(sev) typename orc_shim::Pick<orc_shim::sev>::type()
#define VLOG(n) orc_shim::NullStream()
#define CHECK(c) (c) ? (void)0 : orc_shim::Voidify() & orc_shim::FatalStream() << "Check failed: " #c " "
#define ORC_CHECK_OP(a, b, op) ((a)op(b)) ? (void)0 : orc_shim::Voidify() & orc_shim::FatalStream() << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) ORC_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) ORC_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) ORC_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) ORC_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) ORC_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) ORC_CHECK_OP(a, b, >=)
namespace google { inline void InitGoogleLogging(const char *) {} }
#endif
