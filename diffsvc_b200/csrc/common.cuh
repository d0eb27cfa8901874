// Shared helpers for libdsvc (error plumbing, launch accounting, device buffers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/dsvc.h"

namespace dsvc {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define DSVC_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      dsvc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DSVC_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define DSVC_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != DSVC_OK) return _r; \
  } while (0)

#define DSVC_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      dsvc::set_error(__VA_ARGS__);  \
      return DSVC_EINVAL;            \
    }                                \
  } while (0)

// every kernel launch of the library goes through this so that bench.py can report gpu_launches
#define DSVC_LAUNCH_CHECK()                                                              \
  do {                                                                                    \
    dsvc::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaPeekAtLastError();                                               \
    if (_e != cudaSuccess) {                                                              \
      dsvc::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DSVC_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

int require_device();  // DSVC_OK if an sm_100 device is current, DSVC_ENODEVICE otherwise

// ---- owning device buffer -----------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  // grow-only allocation
  int reserve(size_t n) {
    if (n <= bytes) return DSVC_OK;
    release();
    DSVC_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return DSVC_OK;
  }
  int upload(const void* host, size_t n, cudaStream_t s) {
    DSVC_TRY(reserve(n));
    DSVC_CUDA(cudaMemcpyAsync(p, host, n, cudaMemcpyHostToDevice, s));
    return DSVC_OK;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Raise a kernel's dynamic shared-memory limit.  Function attributes belong to the device (context), so the
// "already done" note is kept per device: one process may drive several GPUs (handles are created on the weight's device).
template <auto Kern>
inline int ensure_dyn_smem(int bytes) {
  static std::atomic<int> done[64];
  int dev = 0;
  DSVC_CUDA(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  if (!tracked || done[dev].load(std::memory_order_relaxed) < bytes) {
    DSVC_CUDA(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (tracked) done[dev].store(bytes, std::memory_order_relaxed);
  }
  return DSVC_OK;
}

// ---- device math that must not be contracted into FMAs (bit-faithful to the reference's
//      separately-rounded tensor ops) -------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float mishf_(float x) {
  // x * tanh(softplus(x)), softplus threshold 20 as torch.nn.functional.softplus
  float sp = x > 20.0f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float lrelu_(float x, float slope) { return x > 0.0f ? x : x * slope; }

// ---- Philox4x32-10 counter-based generator (library's own N(0,1) stream) -------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// two independent N(0,1) from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = ((float)a + 1.0f) * 2.3283064e-10f;  // (0,1]
  float u2 = (float)b * 2.3283064e-10f;
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}
// N(0,1) for linear element index `idx` of stream (seed, stream_id)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream_id, uint64_t idx) {
  uint4 ctr = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), stream_id, 0x5eedu);
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t sel = (uint32_t)(idx & 3);
  float2 n = (sel < 2) ? box_muller(r.x, r.y) : box_muller(r.z, r.w);
  return (sel & 1) ? n.y : n.x;
}
// four independent N(0,1) for vector index `idx4` of stream (seed, stream_id): one Philox call, two Box-Muller pairs
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t stream_id, uint64_t idx4) {
  uint4 ctr = make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), stream_id, 0x5eed4u);
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream_id, uint64_t idx) {
  uint4 ctr = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), stream_id, 0xa11ceu);
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return (float)(r.x >> 8) * (1.0f / 16777216.0f);  // [0,1)
}

}  // namespace dsvc
