// Split-K variant of the dilated-conv contraction for grids that cannot fill the GPU (one clip: 42 tiles of
// 128 frames x 128 channels on 148 SMs).
//
// The three taps of the k=3 conv are three independent K-ranges of 384.  A cluster of three CTAs owns one output
// tile; CTA r runs the tcgen05 main loop of tc_gemm.cuh for tap r only (6 K-stages instead of 18, on 126 SMs
// instead of 84 -- the main loop is shared-memory-bandwidth bound per SM, so spreading it is what shortens it),
// parks its fp32 partial tile in an L2-resident scratch slab, and after a cluster barrier each CTA reduces and
// finishes ONE THIRD OF THE ROWS of the tile: 3 partial loads + the usual fused epilogue functor per element, all
// coalesced (lane <-> 4 consecutive channels).  No DSMEM traffic; the only cross-CTA communication is the slab
// (2 x 64 KB read per CTA) and the barrier.
//
//   phase 0  main loop (TMA -> smem -> tcgen05.mma -> TMEM), identical to tc_gemm_kernel<.,128,1> with taps = 1
//   phase 1  TMEM -> registers (hi*hi + lo*hi  +  hi*lo) -> smem transpose -> slab[tile][r][row][col]
//   barrier.cluster (release / acquire: the slab writes of all three CTAs are visible)
//   phase 2  rows [43r, 43r+43): v = slab[.][0] + slab[.][1] + slab[.][2]; Epi::apply_pair
#pragma once
#include "tc_gemm.cuh"

namespace dsvc {

constexpr int SK_BN = 128;
constexpr int SK_SPLIT = 3;

template <class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_splitk_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const typename Epi::Params ep, float* __restrict__ slab, int T, int K, int N, int dil) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  static_assert(Epi::kPair, "split-K is wired for the gate|filter pair epilogue of the dilated conv");
  constexpr int BN = SK_BN;
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  auto tile_a = [&](int s, int lo) { return smem_base + (uint32_t)s * Cfg::STAGE + (uint32_t)lo * TC_A_TILE; };
  auto tile_b = [&](int s, int lo) { return smem_base + (uint32_t)s * Cfg::STAGE + 2u * TC_A_TILE + (uint32_t)lo * Cfg::B_TILE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (int)cluster_ctarank();                 // tap owned by this CTA (cluster = 3 consecutive blockIdx.y)
  const int nt = (int)blockIdx.y / SK_SPLIT;            // channel tile
  const int m0 = blockIdx.x * TC_BM, b = blockIdx.z;
  const int kblocks = K / TC_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
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

  auto load_b = [&](int kb, int s) {
    const int row = r * N + nt * BN;                    // weight rows of tap r, channel tile nt ([64 gate | 64 filter])
    tma_load_2d(&tmBh, full_bar(s), tile_b(s, 0), kb * TC_BK, row);
    tma_load_2d(&tmBl, full_bar(s), tile_b(s, 1), kb * TC_BK, row);
  };
  auto load_a = [&](int kb, int s) {
    const int frame = m0 + (r - 1) * dil;               // tap r of the centred k=3 conv; out-of-range frames read as zero
    tma_load_3d(&tmAh, full_bar(s), tile_a(s, 0), kb * TC_BK, frame, b);
    tma_load_3d(&tmAl, full_bar(s), tile_a(s, 1), kb * TC_BK, frame, b);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    const int pre = kblocks < STAGES ? kblocks : STAGES;
    if (elect_one_sync()) {
      for (int it = 0; it < pre; ++it) {
        mbar_expect_tx(full_bar(it), Cfg::STAGE);
        load_b(it, it);                                 // weights: constants, requested before the dependency wait
      }
    }
    __syncwarp();
    pdl_wait();                                         // the activation planes were written by the previous kernel
    if (elect_one_sync()) {
      for (int it = 0; it < pre; ++it) load_a(it, it);
    }
    __syncwarp();
    for (int it = pre; it < kblocks; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(full_bar(s), Cfg::STAGE);
        load_a(it, s);
        load_b(it, s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: xh*[wh;wl] as one N=2*BN MMA, then xl*wh into the first BN columns =====
    const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
    const uint32_t idesc2 = umma_idesc_f16(TC_BM, 2 * BN);
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ah = umma_desc_sw128(tile_a(s, 0)), al = umma_desc_sw128(tile_a(s, 1));
        const uint64_t bh = umma_desc_sw128(tile_b(s, 0));
#pragma unroll
        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
          const uint64_t koff = (uint64_t)((k4 * 32) >> 4);
          const uint32_t acc = (it > 0 || k4 > 0) ? 1u : 0u;
          umma_f16(tmem_base, ah + koff, bh + koff, idesc2, acc);
          umma_f16(tmem_base, al + koff, bh + koff, idesc, 1u);
        }
        umma_commit(empty_bar(s));
        if (it == kblocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  }
  pdl_wait();   // every warp: what follows overwrites the slab / the output plane that earlier kernels read

  // ===== phase 1: partial tile -> slab (row-major, coalesced) =====
  const int m_tiles = (int)gridDim.x, n_tiles = (int)gridDim.y / SK_SPLIT;
  const size_t tile_id = ((size_t)b * m_tiles + blockIdx.x) * n_tiles + nt;
  float* my_slab = slab + (tile_id * SK_SPLIT + r) * (size_t)(TC_BM * BN);
  {
    constexpr int CW = BN / 4;                    // columns staged by this warp (32)
    constexpr int STG_LD = BN + 4;
    const int q = warp & 3, cg = warp >> 2;
    mbar_wait(tmem_full_bar, 0u);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + (size_t)q * 32 * STG_LD;
    {
      float v[32], v2[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CW);
      tmem_ld_cols<32>(taddr, v);
      tmem_ld_cols<32>(taddr + (uint32_t)BN, v2);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stage + lane * STG_LD + cg * CW + j * 4) =
            make_float4(v[4 * j] + v2[4 * j], v[4 * j + 1] + v2[4 * j + 1], v[4 * j + 2] + v2[4 * j + 2], v[4 * j + 3] + v2[4 * j + 3]);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");   // the 4 warps of this TMEM lane quarter
    // this warp copies rows cg*8 .. cg*8+7 of the quarter: one 512-byte row per instruction
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = q * 32 + cg * 8 + i;
      const float4 x = *reinterpret_cast<const float4*>(stage + (size_t)(cg * 8 + i) * STG_LD + lane * 4);
      *reinterpret_cast<float4*>(my_slab + (size_t)row * BN + lane * 4) = x;
    }
  }
  tc_fence_before();
  cluster_sync_all();   // release / acquire at cluster scope: all three partial tiles are visible

  // ===== phase 2: this CTA finishes rows [lo, hi) of the tile =====
  {
    constexpr int RPC = (TC_BM + SK_SPLIT - 1) / SK_SPLIT;     // 43
    const int lo = r * RPC, hi = (lo + RPC < TC_BM) ? lo + RPC : TC_BM;
    const float* s0 = slab + (tile_id * SK_SPLIT) * (size_t)(TC_BM * BN);
    const int ch = lane & 15, rsub = lane >> 4;                // 16 float4 gate chunks per row, two rows per warp pass
    const int c0 = nt * (BN / 2) + 4 * ch;                     // output (pair) channel of this lane
    const EpiCol cc = Epi::col(ep, c0);
    for (int row = lo + 2 * warp + rsub; row < hi; row += 2 * (TC_THREADS / 32)) {
      const int p = m0 + row;
      if (p >= T) continue;
      const EpiPre pre = Epi::pre(ep, b, p, c0);
      float g[4] = {0.f, 0.f, 0.f, 0.f}, f[4] = {0.f, 0.f, 0.f, 0.f};
      float4 gv[SK_SPLIT], fv[SK_SPLIT];
#pragma unroll
      for (int k = 0; k < SK_SPLIT; ++k) {
        const float* base = s0 + (size_t)k * (TC_BM * BN) + (size_t)row * BN;
        gv[k] = __ldcg(reinterpret_cast<const float4*>(base + 4 * ch));            // L2: written by the peer CTAs a moment ago
        fv[k] = __ldcg(reinterpret_cast<const float4*>(base + BN / 2 + 4 * ch));
      }
#pragma unroll
      for (int k = 0; k < SK_SPLIT; ++k) {
        g[0] += gv[k].x; g[1] += gv[k].y; g[2] += gv[k].z; g[3] += gv[k].w;
        f[0] += fv[k].x; f[1] += fv[k].y; f[2] += fv[k].z; f[3] += fv[k].w;
      }
      Epi::apply_pair(ep, b, p, c0, g, f, cc, pre);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
#endif
}

// Measured (one 862-frame clip, B200): parity-clean and the kernel itself is 16 % faster run back to back (11.76 ->
// 9.93 us, 130 -> 154 TFLOP/s algorithmic), but the 1000-step sampler does not move (409 vs 406 us per step): with 126
// of 148 SMs holding a 197 KB CTA, the next kernel's CTAs can no longer pre-launch on idle SMs, so the PDL overlap of
// its prologue + weight prefetch -- which is what hid the kernel boundary -- is lost.
// Hence the automatic rule: split only when the 3x larger conv grid AND the next kernel's grid (64-wide tiles) fit
// the GPU together, i.e. short clips / real-time chunks (<= 4 frame tiles in flight).
// DSVC_SPLITK=0 never, =1 whenever the 3x grid alone fits one wave, =2 always.
inline bool tc_splitk_eligible(int B, int T, int N, int taps, int num_sms) {
  const char* ev = getenv("DSVC_SPLITK");          // read per call (graph capture time): tests switch it per handle
  const int env = ev ? atoi(ev) : -1;
  if (env == 0 || taps != SK_SPLIT || N % SK_BN != 0) return false;
  if (env >= 2) return true;
  const long long mt = (long long)ceil_div(T, TC_BM) * B;
  const long long conv_ctas = mt * (N / SK_BN) * SK_SPLIT, next_ctas = mt * (N / 64);
  if (env == 1) return conv_ctas <= num_sms;
  return conv_ctas + next_ctas <= num_sms;
}

inline size_t tc_splitk_slab_bytes(int B, int T, int N) {
  return (size_t)B * ceil_div(T, TC_BM) * (N / SK_BN) * SK_SPLIT * TC_BM * SK_BN * sizeof(float);
}

template <class Epi>
int tc_splitk_launch(const TcGemmMaps& m, const typename Epi::Params& e, float* slab, int B, int T, int K, int N, int dil,
                     cudaStream_t s) {
  DSVC_REQUIRE(K % TC_BK == 0 && N % SK_BN == 0, "tc_splitk_launch: K=%d N=%d", K, N);
  auto kern = tc_splitk_kernel<Epi>;
  DSVC_TRY((ensure_dyn_smem<tc_splitk_kernel<Epi>>(TcCfg<SK_BN>::SMEM)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div(T, TC_BM), (N / SK_BN) * SK_SPLIT, B);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<SK_BN>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 1;
  attr[na].val.clusterDim.y = SK_SPLIT;
  attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DSVC_CUDA(cudaLaunchKernelEx(&cfg, kern, m.a_hi, m.a_lo, m.b_hi, m.b_lo, e, slab, T, K, N, dil));
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
