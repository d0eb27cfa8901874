// tcgen05 / TMA main loop for the WaveNet contractions (sm_100a only).
//
//   D[frame][n] = sum_{tap} sum_{k} A[b][frame + (tap-1)*dil][k] * W[tap][n][k]
//
// A: activation plane, channels-last fp16 (hi, lo) pair [B][T][K]  (K-major)
// W: weight matrix fp16 (hi, lo) pair [taps*N][K]                  (K-major)
// One CTA owns a 128-frame x 128-channel output tile.  Operands are staged by TMA into 128B-swizzled
// shared memory (the shifted tap is just a different TMA frame coordinate; out-of-range frames are
// zero-filled by the TMA unit = the conv's zero padding), multiplied by tcgen05.mma (M=128, N=128,
// K=16 per instruction, fp32 accumulators in TMEM) and read back with tcgen05.ld by four epilogue
// warps that run the fused epilogue functor (epilogues.cuh) straight out of registers.
//
// fp32-class accuracy on fp16 tensor cores: x = xh + xl, w = wh + wl (each fp16), and
//   x*w ~= xh*wh + xl*wh + xh*wl      (3 MMAs into the same fp32 accumulator; the dropped xl*wl
//                                      term is ~2^-22 relative)
// `passes == 1` issues only xh*wh ("fast mode", outside the parity gate).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM lane group = warp % 4).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dsvc {

struct TcGemmMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};
struct TcMaps {
  TcGemmMaps in, skip, head;
  std::vector<TcGemmMaps> dil, out;
};

constexpr int TC_BM = 128;
constexpr int TC_BN = 128;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;            // 16 KB: one [128 rows][64 fp16] tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) {   // ~2 s: a lost arrive must fail loudly, not hang the GPU
      printf("libdsvc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: 8-row groups are 1024 B apart (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address  [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (16 B units) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the kernel -----------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const typename Epi::Params ep, int T, int K, int N, int taps, int dil, int passes) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * TC_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 1);
  auto tile = [&](int s, int which) { return smem_base + (uint32_t)s * TC_STAGE_BYTES + (uint32_t)which * TC_TILE_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN, b = blockIdx.z;
  const int kblocks = K / TC_BK;
  const int total = taps * kblocks;
  const bool three = passes == 3;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (three) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint32_t tx_bytes = three ? TC_STAGE_BYTES : TC_STAGE_BYTES / 2;
      for (int it = 0; it < total; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int tap = it / kblocks, kb = it - tap * kblocks;
        const int frame = m0 + (taps == 3 ? (tap - 1) * dil : 0);
        mbar_expect_tx(full_bar(s), tx_bytes);
        tma_load_3d(&tmAh, full_bar(s), tile(s, 0), kb * TC_BK, frame, b);
        tma_load_2d(&tmBh, full_bar(s), tile(s, 2), kb * TC_BK, tap * N + n0);
        if (three) {
          tma_load_3d(&tmAl, full_bar(s), tile(s, 1), kb * TC_BK, frame, b);
          tma_load_2d(&tmBl, full_bar(s), tile(s, 3), kb * TC_BK, tap * N + n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_f16(TC_BM, TC_BN);
      for (int it = 0; it < total; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t ah = umma_desc_sw128(tile(s, 0)), al = umma_desc_sw128(tile(s, 1));
        const uint64_t bh = umma_desc_sw128(tile(s, 2)), bl = umma_desc_sw128(tile(s, 3));
#pragma unroll
        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
          const uint64_t koff = (uint64_t)((k4 * 32) >> 4);   // +32 B per K=16 step inside the swizzled row
          umma_f16(tmem_base, ah + koff, bh + koff, idesc, (it > 0 || k4 > 0) ? 1u : 0u);
          if (three) {
            umma_f16(tmem_base, al + koff, bh + koff, idesc, 1u);
            umma_f16(tmem_base, ah + koff, bl + koff, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));          // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);           // accumulator complete
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> smem transpose -> fused functor -> coalesced global =====
    // tcgen05.ld hands each thread one accumulator ROW (a frame).  Global tensors are channels-last,
    // so rows are staged through shared memory (the pipeline stages are free once tmem_full fires)
    // and re-read with lane <-> 4 consecutive channels: every global access of the functor is a
    // 512-byte contiguous warp transaction.
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int g = warp & 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16);
    constexpr int STG_LD = TC_BN + 4;                    // padded row: conflict-free float4 writes
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + (size_t)g * 32 * STG_LD;
#pragma unroll 1
    for (int c = 0; c < TC_BN / 32; ++c) {
      float v[32];
      tmem_ld32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * STG_LD + c * 32 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncwarp();
    if constexpr (Epi::kPair) {
      const int l16 = lane & 15;
#pragma unroll 2
      for (int r2 = 0; r2 < 16; ++r2) {
        const int r = 2 * r2 + (lane >> 4);
        const int p = m0 + g * 32 + r;
        const float4 gv = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * l16);
        const float4 fv = *reinterpret_cast<const float4*>(stg + r * STG_LD + 64 + 4 * l16);
        if (p < T) {
          const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
          const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
          Epi::apply_pair(ep, b, p, blockIdx.y * 64 + 4 * l16, gg, ff);
        }
      }
    } else {
      const int n = n0 + 4 * lane;
#pragma unroll 2
      for (int r = 0; r < 32; ++r) {
        const int p = m0 + g * 32 + r;
        const float4 xv = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * lane);
        if (p < T && n < N) {
          const float vv[4] = {xv.x, xv.y, xv.z, xv.w};
          Epi::apply(ep, b, p, n, vv);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_BN) : "memory");
  }
#endif
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int tc_encode_fn(PFN_encodeTiled* out) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    DSVC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled is not available from the installed driver");
      return DSVC_ECUDA;
    }
    fn = (PFN_encodeTiled)p;
  }
  *out = fn;
  return DSVC_OK;
}

// activation plane [B][T][K] fp16, box = {64 channels, 128 frames, 1 item}
static inline int tc_make_a_map(CUtensorMap* m, const __half* base, int B, int T, int K) {
  PFN_encodeTiled enc;
  DSVC_TRY(tc_encode_fn(&enc));
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)T * K * 2};
  cuuint32_t box[3] = {TC_BK, TC_BM, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A [%d][%d][%d]) failed: %d", B, T, K, (int)r); return DSVC_ECUDA; }
  return DSVC_OK;
}
// weight matrix [rows][K] fp16, box = {64, 128 rows}
static inline int tc_make_b_map(CUtensorMap* m, const __half* base, int rows, int K) {
  PFN_encodeTiled enc;
  DSVC_TRY(tc_encode_fn(&enc));
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {TC_BK, TC_BN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B [%d][%d]) failed: %d", rows, K, (int)r); return DSVC_ECUDA; }
  return DSVC_OK;
}

template <class Epi>
int tc_launch(const TcGemmMaps& m, const typename Epi::Params& e, int B, int T, int K, int N, int taps, int dil, int passes,
              cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    DSVC_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  DSVC_REQUIRE(K % TC_BK == 0, "tc_launch: K=%d must be a multiple of %d", K, TC_BK);
  dim3 grid(ceil_div(T, TC_BM), ceil_div(N, TC_BN), B);
  tc_gemm_kernel<Epi><<<grid, 256, TC_SMEM_BYTES, s>>>(m.a_hi, m.a_lo, m.b_hi, m.b_lo, e, T, K, N, taps, dil, passes);
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
