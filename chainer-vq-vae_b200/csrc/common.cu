#include "common.cuh"
#include <string.h>
#include <atomic>

namespace vqw {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
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
