// Persistent "one DiffNet evaluation" kernel for the launch-bound regime (single clips, real-time chunks).
//
// When a clip is short enough that every contraction of the evaluation fits in one wave of 64-wide
// tiles (<= 148 CTAs), the 2L+3 = 43 kernels of an evaluation are dominated by what happens BETWEEN them
// (grid completion -> dependent release ~2.5 us each even with PDL, prologue, TMEM alloc).  This kernel
// runs all 43 contractions as PHASES of one cooperative launch: barriers, TMEM and tensor-map state are
// set up once, each CTA owns (at most) one tile per phase, and phases are separated by a grid-wide
// sense barrier in global memory (~1 us).  The main loop and the fused epilogues are the ones of
// tc_gemm.cuh (BN = 64, 2 MMAs per K-step, smem-staged 16-warp epilogue).
//
// The per-phase description (TMA descriptors, shapes, epilogue parameters) is a table in global memory
// built by the host once per sampler call.
#pragma once
#include "tc_gemm.cuh"

namespace dsvc {

enum StepPhaseType : int { PH_INPROJ = 0, PH_GATE = 1, PH_OUTPROJ = 2, PH_SKIPPROJ = 3, PH_HEAD = 4 };

struct alignas(128) StepPhase {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;        // box {64,128,1} activations; box {64,32} weights
  int type, K, N, taps, dil;
  int m_tiles, n_tiles, tiles;               // tiles = B * m_tiles * n_tiles  (<= gridDim.x)
  union EpiU {
    EpiInProj::Params inproj;
    EpiGate::Params gate;
    EpiOutProj::Params outproj;
    EpiSkipProj::Params skip;
    EpiHead::Params head;
    EpiU() {}
  } ep;
  StepPhase() {}
};

constexpr int STEP_BN = 64;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// self-resetting sense barrier over all CTAs of the (cooperative) grid
__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned* gen) {
  asm volatile("fence.proxy.async;" ::: "memory");   // this CTA's generic-proxy stores vs. others' later TMA reads
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned g = ld_acquire_u32(gen);
    __threadfence();
    if (atomicAdd(count, 1u) == gridDim.x - 1) {
      atomicExch(count, 0u);
      __threadfence();
      atomicAdd(gen, 1u);
    } else {
      const long long t0 = clock64();
      while (ld_acquire_u32(gen) == g) {
        if (clock64() - t0 > 4000000000ll) { printf("libdsvc: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
      }
    }
    __threadfence();
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_step_kernel(const StepPhase* __restrict__ phases, int nphases, int T, int passes, unsigned* gbar) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  constexpr int BN = STEP_BN;
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
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
  const bool three = passes == 3;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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

  uint32_t git = 0;     // pipeline iterations issued so far by this CTA (same count in producer and MMA warp)
  uint32_t ntile = 0;   // tiles this CTA has completed (tmem_full parity)

  for (int p = 0; p < nphases; ++p) {
    const StepPhase& ph = phases[p];
    const int tile = blockIdx.x;
    if (tile < ph.tiles) {
      const int ny = tile % ph.n_tiles;
      const int mt = (tile / ph.n_tiles) % ph.m_tiles;
      const int b = tile / (ph.n_tiles * ph.m_tiles);
      const int m0 = mt * TC_BM, n0 = ny * BN;
      const int K = ph.K, N = ph.N, taps = ph.taps, dil = ph.dil;
      const int kblocks = K / TC_BK;
      const int total = taps * kblocks;
      const bool pair = ph.type == PH_GATE;

      if (warp == 0) {
        // ===== TMA producer =====
        const uint32_t tx_bytes = three ? Cfg::STAGE : Cfg::STAGE / 2;
        for (int it = 0; it < total; ++it) {
          const uint32_t g = git + (uint32_t)it;
          const int s = (int)(g % STAGES);
          const uint32_t par = (g / STAGES) & 1u;
          mbar_wait(empty_bar(s), par ^ 1u);
          if (elect_one_sync()) {
            const int tap = it / kblocks, kb = it - tap * kblocks;
            const int frame = m0 + (taps == 3 ? (tap - 1) * dil : 0);
            mbar_expect_tx(full_bar(s), tx_bytes);
            tma_load_3d(&ph.a_hi, full_bar(s), tile_a(s, 0), kb * TC_BK, frame, b);
            if (three) tma_load_3d(&ph.a_lo, full_bar(s), tile_a(s, 1), kb * TC_BK, frame, b);
            const int r0 = pair ? tap * N + (ny >> 1) * 128 + (ny & 1) * 32 : tap * N + n0;
            const int r1 = pair ? r0 + 64 : r0 + 32;
            tma_load_2d(&ph.b_hi, full_bar(s), tile_b(s, 0), kb * TC_BK, r0);
            tma_load_2d(&ph.b_hi, full_bar(s), tile_b(s, 0) + 32u * 128u, kb * TC_BK, r1);
            if (three) {
              tma_load_2d(&ph.b_lo, full_bar(s), tile_b(s, 1), kb * TC_BK, r0);
              tma_load_2d(&ph.b_lo, full_bar(s), tile_b(s, 1) + 32u * 128u, kb * TC_BK, r1);
            }
          }
          __syncwarp();
        }
      } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
        const uint32_t idesc2 = umma_idesc_f16(TC_BM, 2 * BN);
        for (int it = 0; it < total; ++it) {
          const uint32_t g = git + (uint32_t)it;
          const int s = (int)(g % STAGES);
          const uint32_t par = (g / STAGES) & 1u;
          mbar_wait(full_bar(s), par);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t ah = umma_desc_sw128(tile_a(s, 0)), al = umma_desc_sw128(tile_a(s, 1));
            const uint64_t bh = umma_desc_sw128(tile_b(s, 0));
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
              const uint64_t koff = (uint64_t)((k4 * 32) >> 4);
              const uint32_t acc = (it > 0 || k4 > 0) ? 1u : 0u;
              if (three) {
                umma_f16(tmem_base, ah + koff, bh + koff, idesc2, acc);
                umma_f16(tmem_base, al + koff, bh + koff, idesc, 1u);
              } else {
                umma_f16(tmem_base, ah + koff, bh + koff, idesc, acc);
              }
            }
            umma_commit(empty_bar(s));
            if (it == total - 1) umma_commit(tmem_full_bar);
          }
          __syncwarp();
        }
      }
      git += (uint32_t)total;

      // ===== epilogue (all 16 warps) =====
      const uint32_t tpar = ntile & 1u;
#ifdef DSVC_TIMELINE
      const long long tl0 = 0;
#define STEP_EPI(EPI, FIELD) tc_epilogue<EPI, BN>(ph.ep.FIELD, smem_raw, smem_base, tmem_base, tmem_full_bar, tpar, T, N, m0, ny, b, warp, lane, three, tl0)
#else
#define STEP_EPI(EPI, FIELD) tc_epilogue<EPI, BN>(ph.ep.FIELD, smem_raw, smem_base, tmem_base, tmem_full_bar, tpar, T, N, m0, ny, b, warp, lane, three)
#endif
      switch (ph.type) {
        case PH_INPROJ:   STEP_EPI(EpiInProj, inproj); break;
        case PH_GATE:     STEP_EPI(EpiGate, gate); break;
        case PH_OUTPROJ:  STEP_EPI(EpiOutProj, outproj); break;
        case PH_SKIPPROJ: STEP_EPI(EpiSkipProj, skip); break;
        default:          STEP_EPI(EpiHead, head); break;
      }
#undef STEP_EPI
      ntile += 1;
      tc_fence_before();   // TMEM reads of this tile are ordered before the next tile's MMAs (after the barrier)
    }
    if (p + 1 < nphases) {
      grid_barrier(gbar, gbar + 1);
      tc_fence_after();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
#endif
}

}  // namespace dsvc
