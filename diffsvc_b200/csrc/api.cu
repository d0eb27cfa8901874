// libdsvc: process-wide plumbing of the C-ABI (errors, device probe, launch accounting).
#include "common.cuh"

namespace dsvc {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  // cudaGetDeviceProperties costs tens of milliseconds: probe each device once (the answer cannot change)
  static std::atomic<int> verdict[64];   // 0 unknown, 1 ok, 2 wrong architecture
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && verdict[dev].load(std::memory_order_relaxed) == 1) return DSVC_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: libdsvc has no CPU fallback (%s)", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    return DSVC_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("cannot query the current CUDA device");
    return DSVC_ENODEVICE;
  }
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libdsvc is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    return DSVC_ENODEVICE;
  }
  if (dev >= 0 && dev < 64) verdict[dev].store(1, std::memory_order_relaxed);
  return DSVC_OK;
}

}  // namespace dsvc

extern "C" {

const char* dsvc_version(void) { return "diffsvc-b200 0.1.0 (sm_100a)"; }
#ifndef DSVC_ABI_HASH
#define DSVC_ABI_HASH 0u   /* built without the in-tree build script: the binding then skips the check */
#endif
uint32_t dsvc_abi(void) { return (uint32_t)DSVC_ABI_HASH; }
const char* dsvc_last_error(void) { return dsvc::g_err; }

int dsvc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
  }
  return ok;
}

uint64_t dsvc_launch_count(void) { return dsvc::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
