// Error channel and version of libnerf_b200.so.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace nerf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
constexpr int kMaxDevices = 64, kSlots = 8;
std::mutex g_dev_mutex;
int g_sms[kMaxDevices];                    // 0 = not queried yet
bool g_used[kMaxDevices][kSlots];
}  // namespace

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    cudaGetLastError();
    return 148;
  }
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (g_sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
    g_sms[dev] = n > kMaxSMs ? kMaxSMs : n;
  }
  return g_sms[dev];
}

bool device_first_use(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || slot < 0 || slot >= kSlots) {
    cudaGetLastError();
    return true;                           // unknown device: redo the setup every time (harmless)
  }
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  const bool first = !g_used[dev][slot];
  g_used[dev][slot] = true;
  return first;
}
}  // namespace nerf

extern "C" {
const char* nerf_version(void) { return "nerf_b200 0.1 (sm_100a)"; }
const char* nerf_last_error(void) { return nerf::g_err; }
}
