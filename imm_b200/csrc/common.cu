#include "common.cuh"
#include <stdarg.h>
#include <string.h>

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

// ---- host utility: CRC-32C (Castagnoli), slicing-by-8.  Used by the TensorBundle checkpoint writer/reader
// (imm_b200/utils/tf_checkpoint.py) for block trailers and per-tensor checksums. ----
namespace {
struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
}  // namespace

extern "C" uint32_t immb_crc32c(const void* data, size_t n, uint32_t crc) {
  static const Crc32cTables T;
  const unsigned char* p = (const unsigned char*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
        T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}
