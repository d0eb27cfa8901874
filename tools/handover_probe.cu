// Hand-over latency probe (developer tool, not part of libdsvc).
//
// Question: inside ONE persistent kernel, how long does it take to hand an activation tile from the CTAs that
// produced it (generic-proxy st.global from their epilogue) to the CTAs that consume it as a TMA operand load,
// when the dependency is a per-frame-tile completion counter instead of a kernel boundary / grid barrier?
//
//   producer:  st.global (16 KB per CTA) -> bar.sync -> [fence] -> red.add(cnt[m])
//   consumer:  poll ld.acquire(cnt[m-1..m+1]) -> [fence.proxy.async] -> cp.async.bulk.tensor (4 stages x 32 KB) -> mbarrier
//
// Geometry = the benchmark clip: 7 frame tiles x 12 channel tiles = 84 CTAs x 512 threads, planes [896][384] fp16
// (hi, lo), ping-ponged by phase parity, dependencies on frame tiles m-1, m, m+1 (the dilated conv's halo).
// Every consumer checks that the tile it received holds the previous phase's marker (visibility through the async
// proxy), mismatches are counted.
//
// mode bits: 1 = every producer thread does __threadfence + fence.proxy.async before the CTA barrier (round-1 style)
//                (0: only thread 0 fences after the barrier)
//            2 = consumer issues fence.proxy.async after the acquire
//            4 = sense-reversing grid barrier between phases instead of the counters (round-1 tc_step style)
//            8 = thread 0 uses fence.acq_rel.gpu + red.relaxed instead of __threadfence + atomicAdd
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/handover_probe tools/handover_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int T = 862, TP = 896, K = 384, BM = 128, BK = 64, MT = 7, NT = 12, THREADS = 512, STAGES = 4;
constexpr int A_TILE = BM * BK * 2;          // 16 KB
constexpr int STAGE = 2 * A_TILE;            // hi + lo
constexpr int SMEM = STAGES * STAGE + 1024 + 256 + 64 * 1024;   // + padding: one CTA per SM like the product kernels

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct Probe {
  __half* plane[2][2];     // [parity][hi/lo]  [TP][K]
  unsigned* cnt;           // [MT] (one 128-byte line each) + grid barrier words at [MT*32], [MT*32+32]
  unsigned* errors;
  long long* stamps;       // [grid][8] accumulated cycles
};

__global__ void __launch_bounds__(THREADS, 1)
probe_kernel(const __grid_constant__ CUtensorMap m0h, const __grid_constant__ CUtensorMap m0l,
             const __grid_constant__ CUtensorMap m1h, const __grid_constant__ CUtensorMap m1l, Probe pr, int phases, int mode,
             int spin) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES * STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / NT, ny = blockIdx.x % NT;
  const int row0 = mt * BM;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long acc_poll = 0, acc_first = 0, acc_all = 0, acc_store = 0, acc_signal = 0;
  unsigned* gbar_count = pr.cnt + MT * 32;
  unsigned* gbar_gen = pr.cnt + MT * 32 + 32;
  const long long k0 = clock64();
  for (int p = 1; p <= phases; ++p) {
    const int rp = p & 1;
    const CUtensorMap* mh = rp ? &m1h : &m0h;
    const CUtensorMap* ml = rp ? &m1l : &m0l;
    const long long t0 = clock64();
    long long t1 = t0, t2 = t0;
    if (warp == 0) {
      if (lane == 0) {
        if (!(mode & 4)) {
          const unsigned expected = (unsigned)(NT * (p - 1));
          const int lo = mt > 0 ? mt - 1 : 0, hi = mt + 1 < MT ? mt + 1 : MT - 1;
          for (int m = lo; m <= hi; ++m)
            while (ld_acquire(pr.cnt + m * 32) < expected) { }
        }
        if (mode & 2) asm volatile("fence.proxy.async;" ::: "memory");
        t1 = clock64();
        for (int s = 0; s < STAGES; ++s) {
          mbar_expect_tx(bar0 + 8u * s, STAGE);
          tma_load_3d(mh, bar0 + 8u * s, base + s * STAGE, s * BK, row0 - 1, 0);            // tap -1: reads a halo row of tile m-1
          tma_load_3d(ml, bar0 + 8u * s, base + s * STAGE + A_TILE, s * BK, row0 + 1, 0);   // tap +1: halo row of tile m+1
        }
      }
      __syncwarp();
    }
    const uint32_t par = (uint32_t)(p - 1) & 1u;
    mbar_wait(bar0, par);
    if (threadIdx.x == 0) t2 = clock64();
    for (int s = 1; s < STAGES; ++s) mbar_wait(bar0 + 8u * s, par);
    const long long t3 = clock64();
    // verify: every in-range element equals the previous phase's marker (zero only in out-of-range rows)
    {
      const __half want = __int2half_rn(1 + ((p - 1) % 1000));
      const unsigned short wbits = *reinterpret_cast<const unsigned short*>(&want);
      unsigned bad = 0;
      const uint4* sm = reinterpret_cast<const uint4*>(smem_raw + (base - smem_u32(smem_raw)));
      for (int i = threadIdx.x; i < STAGES * STAGE / 16; i += THREADS) {
        const uint4 v = sm[i];
        const int tile = i / (A_TILE / 16);             // even: hi plane (rows row0-1 ..), odd: lo plane (rows row0+1 ..)
        const int r = (i % (A_TILE / 16)) / 8;          // row inside the tile (128-byte rows; the swizzle permutes chunks within a row)
        const int grow = row0 + r + ((tile & 1) ? 1 : -1);
        const bool inr = grow >= 0 && grow < T;
        const unsigned w = inr ? (((unsigned)wbits << 16) | wbits) : 0u;
        bad += (v.x != w) + (v.y != w) + (v.z != w) + (v.w != w);
      }
      if (bad) atomicAdd(pr.errors, bad);
    }
    // optional stand-in for the main loop + epilogue math
    if (spin > 0) { const long long s0 = clock64(); while (clock64() - s0 < spin) { } }
    __syncthreads();   // everybody is done reading the ring
    const long long t4 = clock64();
    // producer: this CTA's [128 rows x 32 channels] slice of both planes of the other parity
    {
      const __half mk = __int2half_rn(1 + (p % 1000));
      const unsigned short mb = *reinterpret_cast<const unsigned short*>(&mk);
      const unsigned w = ((unsigned)mb << 16) | mb;
      const int r = threadIdx.x >> 2, c = ny * 32 + (threadIdx.x & 3) * 8;
      const int grow = row0 + r;
      if (grow < T) {
        *reinterpret_cast<uint4*>(pr.plane[rp ^ 1][0] + (size_t)grow * K + c) = make_uint4(w, w, w, w);
        *reinterpret_cast<uint4*>(pr.plane[rp ^ 1][1] + (size_t)grow * K + c) = make_uint4(w, w, w, w);
      }
    }
    if (mode & 1) { __threadfence(); asm volatile("fence.proxy.async;" ::: "memory"); }
    __syncthreads();
    const long long t5 = clock64();
    if (threadIdx.x == 0) {
      if (mode & 4) {
        const unsigned g = ld_acquire(gbar_gen);
        __threadfence();
        if (atomicAdd(gbar_count, 1u) == gridDim.x - 1) {
          atomicExch(gbar_count, 0u);
          __threadfence();
          atomicAdd(gbar_gen, 1u);
        } else {
          while (ld_acquire(gbar_gen) == g) { }
        }
        __threadfence();
      } else if (mode & 8) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(pr.cnt + mt * 32) : "memory");
      } else {
        __threadfence();
        atomicAdd(pr.cnt + mt * 32, 1u);
      }
    }
    if (mode & 4) __syncthreads();
    const long long t6 = clock64();
    if (threadIdx.x == 0) {
      acc_poll += t1 - t0; acc_first += t2 - t1; acc_all += t3 - t1; acc_store += t5 - t4; acc_signal += t6 - t5;
    }
  }
  if (threadIdx.x == 0) {
    long long* o = pr.stamps + (size_t)blockIdx.x * 8;
    o[0] = acc_poll; o[1] = acc_first; o[2] = acc_all; o[3] = acc_store; o[4] = acc_signal; o[5] = clock64() - k0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int phases = argc > 1 ? atoi(argv[1]) : 2000;
  CK(cudaSetDevice(0));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  Probe pr{};
  CUtensorMap maps[2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      CK(cudaMalloc(&pr.plane[a][b], (size_t)TP * K * 2));
      cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)T, 1};
      cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)T * K * 2};
      cuuint32_t box[3] = {BK, BM, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&maps[a][b], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, pr.plane[a][b], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
  CK(cudaMalloc(&pr.cnt, (MT * 32 + 64) * sizeof(unsigned)));
  CK(cudaMalloc(&pr.errors, sizeof(unsigned)));
  CK(cudaMalloc(&pr.stamps, (size_t)MT * NT * 8 * sizeof(long long)));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  int clk = 0;
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("handover probe: %d CTAs x %d threads, %d phases, smem %d B, SM clock (max) %d kHz\n", MT * NT, THREADS, phases, SMEM, clk);
  const int modes[] = {0, 2, 8, 10, 1, 3, 4, 6};
  const int spins[] = {0, 8000};
  for (int sp : spins)
  for (int mode : modes) {
    // phase 1 reads parity 1, expecting marker(0) = 1
    std::vector<__half> init((size_t)TP * K, __float2half(1.0f));
    CK(cudaMemcpy(pr.plane[1][0], init.data(), init.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pr.plane[1][1], init.data(), init.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(pr.plane[0][0], 0, (size_t)TP * K * 2));
    CK(cudaMemset(pr.plane[0][1], 0, (size_t)TP * K * 2));
    CK(cudaMemset(pr.cnt, 0, (MT * 32 + 64) * sizeof(unsigned)));
    CK(cudaMemset(pr.errors, 0, sizeof(unsigned)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int ph = phases, md = mode, spn = sp;
    void* args[] = {&maps[0][0], &maps[0][1], &maps[1][0], &maps[1][1], &pr, &ph, &md, &spn};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((void*)probe_kernel, dim3(MT * NT), dim3(THREADS), args, SMEM, 0));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned errs = 0;
    CK(cudaMemcpy(&errs, pr.errors, sizeof(unsigned), cudaMemcpyDeviceToHost));
    std::vector<long long> st((size_t)MT * NT * 8);
    CK(cudaMemcpy(st.data(), pr.stamps, st.size() * 8, cudaMemcpyDeviceToHost));
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < MT * NT; ++c) for (int j = 0; j < 6; ++j) a[j] += (double)st[(size_t)c * 8 + j] / (MT * NT);
    printf("mode %2d spin %5d: %8.3f us/phase | cycles/phase %7.0f | poll-wait %6.0f | flags->first stage %6.0f | ->all 4 stages %6.0f | "
           "stores+fence+bar %6.0f | signal %6.0f | mismatches %u\n",
           mode, sp, 1e3 * ms / phases, a[5] / phases, a[0] / phases, a[1] / phases, a[2] / phases, a[3] / phases, a[4] / phases, errs);
  }
  return 0;
}
