#include "common.cuh"
#include <stdarg.h>

namespace immb {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launch_count{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace immb

extern "C" int immb_version(void) { return IMMB_VERSION; }
extern "C" const char* immb_last_error(void) { return immb::g_last_error; }
extern "C" int64_t immb_launch_count(void) { return (int64_t)immb::g_launch_count.load(); }
