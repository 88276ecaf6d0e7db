// tvo_b200 — C-ABI housekeeping: error reporting, launch counter, device info.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace tvo {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tvo

extern "C" {
int tvo_abi_version(void) { return TVO_ABI_VERSION; }
const char *tvo_last_error(void) { return tvo::g_err; }
int64_t tvo_launch_count(void) { return (int64_t)tvo::g_launches.load(); }
void tvo_reset_launch_count(void) { tvo::g_launches.store(0); }
int tvo_device_sm_count(int *out) {
  if (!out) return tvo::set_error(TVO_EINVAL, "device_sm_count: null pointer");
  int dev = 0, n = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return TVO_OK;
}
}
