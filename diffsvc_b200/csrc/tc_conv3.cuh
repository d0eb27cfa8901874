// tcgen05 dilated-conv (k=3) kernel with halo re-use of the activation tile across the three taps.
//
// The generic kernel (tc_gemm.cuh) re-loads the [128 x 64] activation tile once per tap at a shifted
// frame coordinate: per K-block it moves 3 x 32 KB of activations + 3 weight tiles, and its mainloop
// is bound by operand delivery (~64 B/clk/SM from L2), not by the MMA floor.  Here ONE [144 x 64]
// tile (frames m0-8 .. m0+135, hi and lo) is loaded per K-block and the three taps are three UMMA
// descriptors into it at row offsets 8 + (tap-1)*dil (dil <= 8): activation traffic drops 3x -> 1.125x.
//
// A descriptor may start at a row that is not a multiple of 8, i.e. inside a 1024-byte swizzle atom:
// measured on B200, the UMMA unit applies the 128B swizzle XOR on absolute shared-memory address bits
// (like TMA), so the plain start address works and the "matrix base offset" field must stay 0
// (setting it to (addr >> 7) & 7 gives wrong results).
//
// STATUS: correct, but measured ~8 % SLOWER than the generic kernel (the mainloop is bound by the
// operand-A read of each tcgen05.mma, ~64-72 cycles per instruction, not by L2->SM bytes), so it is
// opt-in (DSVC_CONV_HALO=1) and kept as a documented experiment.
//
// Two mbarrier rings with their own producer warps: A (3 slots x 36 KB, one per K-block, warp 3) and
// B (weight tiles, one per (K-block, tap), warp 0).
#pragma once
#include "tc_gemm.cuh"

namespace dsvc {

constexpr int TC3_HALO = 8;
constexpr int TC3_AROWS = TC_BM + 2 * TC3_HALO;          // 144 frames
constexpr int TC3_A_TILE = TC3_AROWS * TC_BK * 2;        // 18 KB (18 swizzle atoms)
constexpr int TC3_A_SLOT = 2 * TC3_A_TILE;               // hi + lo
constexpr int TC3_NA = 3;

template <int BN> struct Tc3Cfg {
  static constexpr int B_TILE = BN * TC_BK * 2;
  static constexpr int B_SLOT = 2 * B_TILE;               // hi + lo
  static constexpr int NB = (BN == 64) ? 5 : 3;
  static constexpr int SMEM = TC3_NA * TC3_A_SLOT + NB * B_SLOT + 1024 + 256;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_off(uint32_t saddr) {
  return umma_desc_sw128(saddr) | ((uint64_t)((saddr >> 7) & 7u) << 49);   // matrix base offset
}

template <class Epi, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const typename Epi::Params ep, int T, int K, int N, int dil, int passes, int bo_mode) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  using Cfg = Tc3Cfg<BN>;
  constexpr int NB = Cfg::NB;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + TC3_NA * TC3_A_SLOT;
  const uint32_t bar_base = b_base + NB * Cfg::B_SLOT;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (TC3_NA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * TC3_NA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * TC3_NA + NB + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * TC3_NA + 2 * NB);
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  auto a_tile = [&](int s, int lo) { return smem_base + (uint32_t)s * TC3_A_SLOT + (uint32_t)lo * TC3_A_TILE; };
  auto b_tile = [&](int s, int lo) { return b_base + (uint32_t)s * Cfg::B_SLOT + (uint32_t)lo * Cfg::B_TILE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef DSVC_TIMELINE
  const long long tl0 = clock64();
#endif
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int kblocks = K / TC_BK;
  const int total = 3 * kblocks;
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
    for (int s = 0; s < TC3_NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  if (warp == 4) TL_MARK(0);

  // weight tile(s) of iteration it = kb*3 + tap into B slot s
  auto load_b = [&](int it, int s) {
    const int kb = it / 3, tap = it - kb * 3;
    mbar_expect_tx(b_full(s), three ? Cfg::B_SLOT : Cfg::B_TILE);
    if constexpr (BN == 128) {
      const int row = tap * N + n0;
      tma_load_2d(&tmBh, b_full(s), b_tile(s, 0), kb * TC_BK, row);
      if (three) tma_load_2d(&tmBl, b_full(s), b_tile(s, 1), kb * TC_BK, row);
    } else {
      const int r0 = Epi::kPair ? tap * N + (int)(blockIdx.y >> 1) * 128 + (int)(blockIdx.y & 1) * 32 : tap * N + n0;
      const int r1 = Epi::kPair ? r0 + 64 : r0 + 32;
      tma_load_2d(&tmBh, b_full(s), b_tile(s, 0), kb * TC_BK, r0);
      tma_load_2d(&tmBh, b_full(s), b_tile(s, 0) + 32u * 128u, kb * TC_BK, r1);
      if (three) {
        tma_load_2d(&tmBl, b_full(s), b_tile(s, 1), kb * TC_BK, r0);
        tma_load_2d(&tmBl, b_full(s), b_tile(s, 1) + 32u * 128u, kb * TC_BK, r1);
      }
    }
  };
  auto load_a = [&](int kb, int s) {
    mbar_expect_tx(a_full(s), three ? TC3_A_SLOT : TC3_A_TILE);
    tma_load_3d(&tmAh, a_full(s), a_tile(s, 0), kb * TC_BK, m0 - TC3_HALO, b);
    if (three) tma_load_3d(&tmAl, a_full(s), a_tile(s, 1), kb * TC_BK, m0 - TC3_HALO, b);
  };

  if (warp == 0) {
    // ===== weight-tile producer (B ring) =====
    const int preb = total < NB ? total : NB;
    if (elect_one_sync()) {
      for (int it = 0; it < preb; ++it) load_b(it, it);        // weights: no dependency on the previous kernel
    }
    __syncwarp();
    for (int it = preb; it < total; ++it) {
      const int s = it % NB;
      mbar_wait(b_empty(s), ((uint32_t)(it / NB) & 1u) ^ 1u);
      if (elect_one_sync()) load_b(it, s);
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===== activation-tile producer (A ring), decoupled from the B ring =====
    pdl_wait();                                                // the plane was written by the previous kernel
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % TC3_NA;
      if (kb >= TC3_NA) mbar_wait(a_empty(s), ((uint32_t)(kb / TC3_NA) & 1u) ^ 1u);
      if (elect_one_sync()) load_a(kb, s);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
    const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
    const uint32_t idesc2 = umma_idesc_f16(TC_BM, 2 * BN);   // [wh ; wl] concatenated along N
    for (int kb = 0; kb < kblocks; ++kb) {
      const int sa = kb % TC3_NA;
      mbar_wait(a_full(sa), (uint32_t)(kb / TC3_NA) & 1u);
      if (kb == 0) TL_MARK(1);
      for (int tap = 0; tap < 3; ++tap) {
        const int it = kb * 3 + tap;
        const int sb = it % NB;
        mbar_wait(b_full(sb), (uint32_t)(it / NB) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t roff = (uint32_t)(TC3_HALO + (tap - 1) * dil) * 128u;
          const uint64_t bh = umma_desc_sw128(b_tile(sb, 0));
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
            const uint32_t aaddr_h = a_tile(sa, 0) + roff + (uint32_t)k4 * 32u;
            const uint32_t aaddr_l = a_tile(sa, 1) + roff + (uint32_t)k4 * 32u;
            const uint64_t ah = bo_mode ? umma_desc_sw128_off(aaddr_h) : umma_desc_sw128(aaddr_h);
            const uint64_t al = bo_mode ? umma_desc_sw128_off(aaddr_l) : umma_desc_sw128(aaddr_l);
            const uint64_t koff = (uint64_t)((k4 * 32) >> 4);
            const uint32_t acc = (it > 0 || k4 > 0) ? 1u : 0u;
            if (three) {
              umma_f16(tmem_base, ah, bh + koff, idesc2, acc);
              umma_f16(tmem_base, al, bh + koff, idesc, 1u);
            } else {
              umma_f16(tmem_base, ah, bh + koff, idesc, acc);
            }
          }
          umma_commit(b_empty(sb));
          if (tap == 2) umma_commit(a_empty(sa));
          if (it == total - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    }
    TL_MARK(2);
  }
  pdl_wait();
#ifdef DSVC_TIMELINE
  tc_epilogue<Epi, BN>(ep, smem_raw, smem_base, tmem_base, tmem_full_bar, 0u, T, N, m0, (int)blockIdx.y, b, warp, lane, three, tl0);
#else
  tc_epilogue<Epi, BN>(ep, smem_raw, smem_base, tmem_base, tmem_full_bar, 0u, T, N, m0, (int)blockIdx.y, b, warp, lane, three);
#endif
  if (warp == 4) TL_MARK(6);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
#endif
}

// activation plane [B][T][K] fp16, box = {64 channels, 144 frames, 1 item}
static inline int tc3_make_a_map(CUtensorMap* m, const __half* base, int B, int T, int K) {
  PFN_encodeTiled enc;
  DSVC_TRY(tc_encode_fn(&enc));
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)T * K * 2};
  cuuint32_t box[3] = {TC_BK, TC3_AROWS, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A144 [%d][%d][%d]) failed: %d", B, T, K, (int)r); return DSVC_ECUDA; }
  return DSVC_OK;
}

inline bool tc_use_halo() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DSVC_CONV_HALO"); v = (e && e[0] == '1') ? 1 : 0; }   // opt-in: measured slower
  return v == 1;
}

template <class Epi, int BN>
int tc3_launch_bn(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                  const typename Epi::Params& e, int B, int T, int K, int N, int dil, int passes, cudaStream_t s) {
  static bool attr_set = false;
  auto kern = tc_conv3_kernel<Epi, BN>;
  if (!attr_set) {
    DSVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc3Cfg<BN>::SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div(T, TC_BM), ceil_div(N, BN), B);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Tc3Cfg<BN>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_use_pdl() ? 1 : 0;
  static int bo_mode = -1;
  // measured: the hardware swizzles on absolute smem address bits; the base-offset field stays 0
  if (bo_mode < 0) { const char* ev = getenv("DSVC_BO_MODE"); bo_mode = ev ? atoi(ev) : 0; }
  DSVC_CUDA(cudaLaunchKernelEx(&cfg, kern, ah, al, bh, bl, e, T, K, N, dil, passes, bo_mode));
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
