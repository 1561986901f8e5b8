// Error channel and version of libnerf_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace nerf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nerf

extern "C" {
const char* nerf_version(void) { return "nerf_b200 0.1 (sm_100a)"; }
const char* nerf_last_error(void) { return nerf::g_err; }
}
