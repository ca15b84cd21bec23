// Process-wide pieces of the C ABI: error text, version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace mt3 {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

std::atomic<uint64_t> g_launch_count{0};

}  // namespace mt3

extern "C" int mt3_abi_version(void) { return MT3_ABI_VERSION; }
extern "C" const char* mt3_last_error(void) { return mt3::last_error().c_str(); }
extern "C" uint64_t mt3_kernel_launch_count(void) { return mt3::g_launch_count.load(); }
