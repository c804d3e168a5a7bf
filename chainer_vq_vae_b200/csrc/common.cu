#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <atomic>

namespace vqw {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
bool debug_sync() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQW_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
char* error_buffer() { return g_err; }
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace vqw

namespace vqw { long long launches(); }
extern "C" int vqw_version(void) { return VQW_VERSION; }
extern "C" long long vqw_launch_count(void) { return vqw::launches(); }
extern "C" const char* vqw_last_error(void) { return vqw::error_buffer(); }
