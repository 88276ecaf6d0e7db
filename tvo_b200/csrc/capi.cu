// tvo_b200 — C-ABI housekeeping: error reporting, launch counter, device info.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

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

// ---- opt-in timing of the dominant kernels (bench.py's roofline): cudaEvent pairs recorded on the
// launching stream around each launch of a timed kernel; tvo_kernel_timing_read synchronises them.
static std::atomic<int> g_timing{0};
struct TimedLaunch {
  int which;
  cudaEvent_t a, b;
};
static std::mutex g_timing_mu;
static std::vector<TimedLaunch> g_timed;

bool kernel_timing_on() { return g_timing.load(std::memory_order_relaxed) != 0; }
void kernel_timing_begin(int which, cudaStream_t st, void **token) {
  *token = nullptr;
  if (!kernel_timing_on()) return;
  TimedLaunch *t = new TimedLaunch;
  t->which = which;
  if (cudaEventCreate(&t->a) != cudaSuccess || cudaEventCreate(&t->b) != cudaSuccess) {
    delete t;
    return;
  }
  cudaEventRecord(t->a, st);
  *token = t;
}
void kernel_timing_end(void *token, cudaStream_t st) {
  if (!token) return;
  TimedLaunch *t = (TimedLaunch *)token;
  cudaEventRecord(t->b, st);
  std::lock_guard<std::mutex> g(g_timing_mu);
  g_timed.push_back(*t);
  delete t;
}

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
void tvo_kernel_timing_enable(int on) {
  tvo::g_timing.store(on ? 1 : 0);
  std::lock_guard<std::mutex> g(tvo::g_timing_mu);
  for (auto &t : tvo::g_timed) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  tvo::g_timed.clear();
}
int tvo_kernel_timing_read(int which, double *total_ms, int64_t *launches) {
  if (!total_ms || !launches) return tvo::set_error(TVO_EINVAL, "kernel_timing_read: null pointer");
  std::lock_guard<std::mutex> g(tvo::g_timing_mu);
  double tot = 0.0;
  int64_t n = 0;
  for (auto &t : tvo::g_timed) {
    if (t.which != which) continue;
    TVO_CUDA(cudaEventSynchronize(t.b));
    float ms = 0.f;
    TVO_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *launches = n;
  return TVO_OK;
}
int tvo_device_sm_count(int *out) {
  if (!out) return tvo::set_error(TVO_EINVAL, "device_sm_count: null pointer");
  int dev = 0, n = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return TVO_OK;
}
}
