// Library queries and error plumbing of librobotgnn.so (include/robot_gnn.h).
#include <stdarg.h>
#include <atomic>
#include <string.h>
#include "common.cuh"

namespace rgnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int num_sms() {
  static std::atomic<int> cache[64];
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 148;
  int n = cache[d].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace rgnn

extern "C" long long rgnn_launch_count(void) { return rgnn::g_launches.load(std::memory_order_relaxed); }
extern "C" int rgnn_version(void) { return RGNN_VERSION; }
extern "C" int rgnn_arch(void) { return 100; }
extern "C" const char* rgnn_last_error(void) { return rgnn::g_err; }

extern "C" int rgnn_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return rgnn::cuda_fail(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    rgnn::set_error("librobotgnn is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return RGNN_E_UNSUPPORTED;
  }
  return 0;
}
