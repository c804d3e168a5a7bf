#include "common.cuh"
#include <string.h>

namespace vqw {
static thread_local char g_err[512] = "";
char* error_buffer() { return g_err; }
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace vqw

extern "C" int vqw_version(void) { return VQW_VERSION; }
extern "C" const char* vqw_last_error(void) { return vqw::error_buffer(); }
