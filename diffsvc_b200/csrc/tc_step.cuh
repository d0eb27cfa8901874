// One launch per denoiser evaluation: the 2L+3 contractions of DiffNet as PHASES of one kernel of resident CTA pairs.
//
// Why it was built: at one clip a WaveNet layer is two kernels whose boundary costs as much as their main loops --
// the dependent grid must drain, half of the next grid's CTAs cannot be pre-launched (96 + 96 CTAs > 148 SMs: they start
// 1.5 us late and pay barrier init / TMEM alloc / cluster sync / cold tensor maps again), and the first operand tiles
// travel L2 -> smem only then.  Here every CTA pair owns one (256-frame tile, channel tile) slot for the whole
// evaluation: barriers, TMEM and tensor maps are set up once, the smem ring and its mbarrier phases run on across the
// phases, the next phase's weight tiles are requested BEFORE its dependency wait, and a phase boundary is
//     epilogue stores -> bar.sync -> fence.acq_rel.gpu -> red.add(flag[frame tile])               (producer side)
//     3 lanes poll flag[ft-1], flag[ft], flag[ft+1] in parallel -> fence.proxy.async -> TMA       (consumer side)
// i.e. a dependency on the THREE neighbouring frame tiles only (a dilated tap reaches at most max-dilation rows into the
// neighbour), not on the whole grid.  The uniform rule "a phase starts at frame tile ft when every tile of every earlier
// phase at ft-1, ft, ft+1 is complete" covers all hazards of the layer chain: Y (out-proj -> next conv, read with a
// halo), Z (conv -> out-proj; overwritten by the next conv only after the out-projections at ft-1..ft+1 are done), and
// the in-place residual / skip / sampler tensors, which are only ever touched by the slot (same CTA, same thread) that
// owns their rows and columns.
//
// The tile math is tc_pair.cuh's (cta_group::2, M = 256, [wh_a ; wl_b] / [wh_b ; wl_a] weight blocks, 3-pass product) and
// the epilogues are the shared functors: results are bit-identical to the per-layer pair kernels of the same tile width.
// The whole plan (tensor maps of every phase, epilogue parameter blocks, phase table: ~22 KB) is ONE __grid_constant__
// kernel parameter, so the per-phase scalars stay in the constant bank exactly as in the per-layer kernels.
//
// Residency: all CTAs must be co-resident (a phase polls flags other CTAs raise).  The launcher sizes the grid from
// cudaOccupancyMaxActiveClusters (a pair then owns ceil(slots / pairs) slots) and a dependency wait that lasts ~4 s traps
// instead of hanging the GPU.  Opt-in (DSVC_STEP=1): measured slower than the per-layer kernels, see DESIGN.md 3.1f.
#pragma once
#include "tc_pair.cuh"

namespace dsvc {

constexpr int STEP_MAXL = 24;
constexpr int STEP_MAP_IN = 12, STEP_MAP_SKIP = 14, STEP_MAP_HEAD = 16, STEP_MAP_LAYER0 = 18;   // weight maps (hi, lo pairs)
constexpr int STEP_MAP_XIN = 0, STEP_MAP_Y = 2, STEP_MAP_Z = 6, STEP_MAP_SP = 8, STEP_MAP_R = 10;   // activation maps (hi, lo)
constexpr int STEP_NMAPS = STEP_MAP_LAYER0 + 4 * STEP_MAXL;

enum StepKind : int { STEP_IN = 0, STEP_GATE = 1, STEP_OUT = 2, STEP_SKIP = 3, STEP_HEAD = 4 };

struct StepPhase {
  int kind, K, taps, dil;
  int N;         // output channels = weight rows per tap
  int n_tiles;   // channel tiles of this phase; slots with ny >= n_tiles sit the phase out
  int a_map, b_map;
  int expect;    // flag value that says "every earlier phase is complete at this frame tile" (0: nothing to wait for)
  int epi;       // index into the parameter array of this kind
};

struct StepPlan {
  CUtensorMap maps[STEP_NMAPS];
  StepPhase phase[2 * STEP_MAXL + 3];
  EpiInProj::Params in;
  EpiGate::Params gate[STEP_MAXL];
  EpiOutProj::Params out[STEP_MAXL];
  EpiSkipProj::Params skip;
  EpiHead::Params head;
  unsigned* flags;   // [2][n_ft][32]: one counter per frame tile on its own 128-byte line, two sets alternating by launch
  unsigned* seq;     // launch sequence number, advanced by CTA 0 when every flag has reached `final`
  int n_phases, n_ft, n_slots, n_maps, T, final;
};
static_assert(sizeof(StepPlan) <= 32000, "the plan must fit the 32 KB kernel-parameter space");

#ifdef DSVC_TIMELINE
#define STEP_TL_ARGS , 0ll, 0
// per (CTA, phase) absolute SM-clock stamps: 0 phase start | 1 dependencies met | 2 first operands landed | 3 MMAs issued |
// 4 accumulator ready | 5 staged | 6 epilogue done | 7 signalled | 8 %globaltimer (ns) at phase start
__device__ long long g_step_tl[160][2 * STEP_MAXL + 3][10];
#define STL(slot) do { g_step_tl[blockIdx.x % 160][ph][slot] = clock64(); } while (0)
#else
#define STEP_TL_ARGS
#define STL(slot) do { } while (0)
#endif

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Thread layout: warps 0..15 = epilogue (TMEM lane quarter = warp % 4, as in tc_epilogue), warp 16 = TMA producer,
// warp 17 = MMA issuer (even CTA).  The three roles walk the same tile sequence (phase-major, the pair's slots inside a
// phase) and meet only through mbarriers, so the producer and the issuer run AHEAD of the epilogue:
//   ring full/empty[STAGES]      producer <-> issuer          (as in tc_pair.cuh; the counters run on across tiles)
//   acc_full[2] / acc_free[2]    issuer -> epilogue -> issuer  two 2*BN-column accumulators in TMEM: tile k+1's MMAs
//                                                              run while tile k drains
// The epilogue stages through its own slab (not the operand ring, which already receives the next tile), and every
// epilogue WARP releases its own stores (fence.acq_rel.gpu + red.add on the frame tile's counter: 16 per CTA and tile).
constexpr int STEP_THREADS = 18 * 32;
constexpr int STEP_SIGNALS = 32;             // per tile: 16 epilogue warps x 2 CTAs

template <int BN> struct StepCfg {
  static constexpr int H = BN / 2;
  static constexpr int STAGE = 2 * TC_A_TILE + BN * TC_BK * 2;
  static constexpr int SLAB = 4 * 32 * (BN + 4) * 4;
  static constexpr int STAGES = (227 * 1024 - 1024 - 256 - SLAB) / STAGE;
  static constexpr int SMEM = STAGES * STAGE + 256 + SLAB + 1024;
  static_assert(STAGES >= 3, "step kernel: at least three operand stages");
};

template <int BN>
__global__ void __launch_bounds__(STEP_THREADS, 1)
tc_step_kernel(const __grid_constant__ StepPlan plan) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  using Cfg = StepCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int H = Cfg::H;
  static_assert(BN == 64 || BN == 128, "step kernel: 64- and 128-wide slots (two 2*BN-column accumulators in TMEM)");
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE;
  const uint32_t slab_base = bar_base + 256u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto acc_free_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto tile_a = [&](int s, int lo) { return smem_base + (uint32_t)s * Cfg::STAGE + (uint32_t)lo * TC_A_TILE; };
  auto tile_p = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE + 2u * TC_A_TILE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const int T = plan.T;
  const int n_all = plan.n_ft * plan.n_slots;   // (frame tile, channel tile) slots; pair p owns slots p, p + pairs, p + 2 pairs, ...

  if (warp == 3) {
    for (int i = lane; i < plan.n_maps; i += 32) asm volatile("prefetch.tensormap [%0];" ::"l"(&plan.maps[i]) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_free_bar(b), STEP_SIGNALS);     // every epilogue warp of both CTAs (used in the even CTA only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(4 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  pdl_wait();                                   // everything below reads what the previous launch wrote (incl. `seq`)
  const unsigned sq = ld_acquire_u32(plan.seq);
  unsigned* const flags = plan.flags + (size_t)(sq & 1u) * plan.n_ft * 32;
  if (blockIdx.x == 0 && (int)threadIdx.x < plan.n_ft)     // the other set was the previous launch's: clean it for the next one
    plan.flags[(size_t)((sq & 1u) ^ 1u) * plan.n_ft * 32 + threadIdx.x * 32] = 0u;

  auto full_bar_leader = [&](int s) { return mapa_cluster(full_bar(s), 0u); };

  if (warp == 16) {
    // ============================== TMA producer ==============================
    uint32_t git = 0;         // ring iterations so far (all tiles): stage = git % STAGES, parity from git / STAGES
    constexpr uint32_t tx_pair = 2u * Cfg::STAGE;
    for (int ph = 0; ph < plan.n_phases; ++ph) {
      const StepPhase& P = plan.phase[ph];
      const int kblocks = P.K / TC_BK, total = P.taps * kblocks;
      const CUtensorMap* tmAh = &plan.maps[P.a_map];
      const CUtensorMap* tmBh = &plan.maps[P.b_map];
      const bool gate = P.kind == STEP_GATE;
      for (int slot = pair; slot < n_all; slot += n_pairs) {
        const int ft = slot / plan.n_slots, ny = slot - ft * plan.n_slots;
        if (ny >= P.n_tiles) continue;
        const int m0 = ft * 2 * TC_BM + (int)rank * TC_BM;
        if (lane == 0) STL(0);
        auto load_p = [&](int it, int s) {
          const int tap = it / kblocks, kb = it - tap * kblocks;
          int ra, rb;
          if (gate) {             // gate|filter packing: a 128-row super-tile = [64 gate rows | 64 filter rows]
            constexpr int per = 128 / BN;
            ra = tap * P.N + (ny / per) * 128 + (ny % per) * H;
            rb = ra + 64;
          } else {
            ra = tap * P.N + ny * BN;
            rb = ra + H;
          }
          const uint32_t bar = full_bar_leader(s);
          const uint32_t p = tile_p(s);
          tma2_load_2d(tmBh, bar, p, kb * TC_BK, rank == 0 ? ra : rb);            // even: [wh_a ; wl_b]   odd: [wh_b ; wl_a]
          tma2_load_2d(tmBh + 1, bar, p + (uint32_t)H * 128u, kb * TC_BK, rank == 0 ? rb : ra);
        };
        auto load_a = [&](int it, int s) {
          const int tap = it / kblocks, kb = it - tap * kblocks;
          const int frame = m0 + (tap - (P.taps >> 1)) * P.dil;
          const uint32_t bar = full_bar_leader(s);
          tma2_load_3d(tmAh, bar, tile_a(s, 0), kb * TC_BK, frame, 0);
          tma2_load_3d(tmAh + 1, bar, tile_a(s, 1), kb * TC_BK, frame, 0);
        };
        const int pre = total < STAGES ? total : STAGES;
        for (int j = 0; j < pre; ++j) {           // weights do not depend on anybody: request them before the flag wait
          const uint32_t g = git + (uint32_t)j;
          const int s = (int)(g % STAGES);
          mbar_wait(empty_bar(s), ((g / STAGES) & 1u) ^ 1u);
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(full_bar(s), tx_pair);
            load_p(j, s);
          }
          __syncwarp();
        }
        if (P.expect > 0) {
          // every earlier phase complete at frame tiles ft-1, ft, ft+1: three lanes poll one counter each
          const int f = ft - 1 + lane;
          bool ok = lane > 2 || f < 0 || f >= plan.n_ft;
          const long long t0 = clock64();
          while (true) {
            if (!ok) ok = ld_acquire_u32(flags + (size_t)f * 32) >= (unsigned)P.expect;
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > 8000000000ll) {   // ~4 s: a lost signal fails loudly instead of hanging the GPU
              if (lane == 0) printf("libdsvc: step kernel dependency wait timed out (block %d phase %d)\n", (int)blockIdx.x, ph);
              __trap();
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores of the producers -> our TMA reads
        }
        if (lane == 0) STL(1);
        if (elect_one_sync()) {
          for (int j = 0; j < pre; ++j) load_a(j, (int)((git + (uint32_t)j) % STAGES));
        }
        __syncwarp();
        for (int j = pre; j < total; ++j) {
          const uint32_t g = git + (uint32_t)j;
          const int s = (int)(g % STAGES);
          mbar_wait(empty_bar(s), ((g / STAGES) & 1u) ^ 1u);
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(full_bar(s), tx_pair);
            load_a(j, s);
            load_p(j, s);
          }
          __syncwarp();
        }
        git += (uint32_t)total;
      }
    }
  } else if (warp == 17) {
    // ============================== MMA issuer: the even CTA, for both ==============================
    if (rank == 0) {
      uint32_t git = 0, k = 0;      // ring iterations, tiles
      const uint32_t idesc_hi = umma_idesc_f16(2 * TC_BM, 2 * BN);
      const uint32_t idesc_lo = umma_idesc_f16(2 * TC_BM, BN);
      for (int ph = 0; ph < plan.n_phases; ++ph) {
        const StepPhase& P = plan.phase[ph];
        const int total = P.taps * (P.K / TC_BK);
        for (int slot = pair; slot < n_all; slot += n_pairs) {
          const int ny = slot % plan.n_slots;
          if (ny >= P.n_tiles) continue;
          const uint32_t buf = k & 1u;
          mbar_wait(acc_free_bar((int)buf), ((k >> 1) & 1u) ^ 1u);     // both CTAs have drained this accumulator
          tc_fence_after();
          const uint32_t acc_addr = tmem_base + buf * (uint32_t)(2 * BN);
          for (int j = 0; j < total; ++j) {
            const uint32_t g = git + (uint32_t)j;
            const int s = (int)(g % STAGES);
            mbar_wait(full_bar(s), (g / STAGES) & 1u);
            if (j == 0 && lane == 0) STL(2);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t ah = umma_desc_sw128(tile_a(s, 0)), al = umma_desc_sw128(tile_a(s, 1));
              const uint64_t pd = umma_desc_sw128(tile_p(s));
#pragma unroll
              for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                const uint64_t koff = (uint64_t)((k4 * 32) >> 4);
                const uint32_t acc = (j > 0 || k4 > 0) ? 1u : 0u;
                umma2_f16(acc_addr, ah + koff, pd + koff, idesc_hi, acc);
                umma2_f16(acc_addr, al + koff, pd + koff, idesc_lo, 1u);
              }
              umma2_commit_both(empty_bar(s));
              if (j == total - 1) umma2_commit_both(acc_full_bar((int)buf));
            }
            __syncwarp();
          }
          if (lane == 0) STL(3);
          git += (uint32_t)total;
          ++k;
        }
      }
    }
  } else if (warp < 16) {
    // ============================== epilogue warps ==============================
    uint32_t k = 0;
    for (int ph = 0; ph < plan.n_phases; ++ph) {
      const StepPhase& P = plan.phase[ph];
      for (int slot = pair; slot < n_all; slot += n_pairs) {
        const int ft = slot / plan.n_slots, ny = slot - ft * plan.n_slots;
        if (ny >= P.n_tiles) continue;
        const int m0 = ft * 2 * TC_BM + (int)rank * TC_BM;
        const uint32_t buf = k & 1u, par = (k >> 1) & 1u;
        const uint32_t acc_addr = tmem_base + buf * (uint32_t)(2 * BN);
        const uint32_t full = acc_full_bar((int)buf);
        const uint32_t free_leader = mapa_cluster(acc_free_bar((int)buf), 0u);
        switch (P.kind) {
          case STEP_IN:
            tc_epilogue<EpiInProj, BN, true>(plan.in, smem_raw, slab_base, acc_addr, full, par, T, P.N, m0, ny, 0, warp, lane, true STEP_TL_ARGS, H, free_leader);
            break;
          case STEP_GATE:
            tc_epilogue<EpiGate, BN, true>(plan.gate[P.epi], smem_raw, slab_base, acc_addr, full, par, T, P.N, m0, ny, 0, warp, lane, true STEP_TL_ARGS, H, free_leader);
            break;
          case STEP_OUT:
            tc_epilogue<EpiOutProj, BN, true>(plan.out[P.epi], smem_raw, slab_base, acc_addr, full, par, T, P.N, m0, ny, 0, warp, lane, true STEP_TL_ARGS, H, free_leader);
            break;
          case STEP_SKIP:
            tc_epilogue<EpiSkipProj, BN, true>(plan.skip, smem_raw, slab_base, acc_addr, full, par, T, P.N, m0, ny, 0, warp, lane, true STEP_TL_ARGS, H, free_leader);
            break;
          default:
            tc_epilogue<EpiHead, BN, true>(plan.head, smem_raw, slab_base, acc_addr, full, par, T, P.N, m0, ny, 0, warp, lane, true STEP_TL_ARGS, H, free_leader);
            break;
        }
#ifdef DSVC_TIMELINE
        if (warp == 4 && lane == 0) {
          g_step_tl[blockIdx.x % 160][ph][4] = g_timeline[blockIdx.x & 1023][4];
          g_step_tl[blockIdx.x % 160][ph][5] = g_timeline[blockIdx.x & 1023][5];
          STL(6);
        }
#endif
        // this warp's stores -> visible, then one more completion on the frame tile's counter
        __syncwarp();
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flags + (size_t)ft * 32) : "memory");
        }
        if (warp == 4 && lane == 0) STL(7);
        ++k;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * BN) : "memory");
  }
  if (blockIdx.x == 0 && warp == 0) {
    // launch epilogue: when every frame tile has seen every phase, nobody reads `seq` or this flag set any more
    const long long t0 = clock64();
    for (int f = lane; f < plan.n_ft; f += 32) {
      while (ld_acquire_u32(flags + (size_t)f * 32) < (unsigned)plan.final) {
        if (clock64() - t0 > 8000000000ll) __trap();
      }
    }
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      *reinterpret_cast<volatile unsigned*>(plan.seq) = sq + 1u;
    }
  }
#endif
}

// Opt-in (DSVC_STEP=1, read per prepare: tests switch it per handle): measured slower than the per-layer kernels at every
// batch size, kept as the parity-tested reference point of the "one launch per evaluation" design.
inline bool tc_step_enabled() {
  const char* e = getenv("DSVC_STEP");
  return e && atoi(e) >= 1;
}

template <int BN>
int tc_step_max_pairs(int* out) {
  DSVC_TRY((ensure_dyn_smem<tc_step_kernel<BN>>(StepCfg<BN>::SMEM)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2, 1, 1);
  cfg.blockDim = dim3(STEP_THREADS);
  cfg.dynamicSmemBytes = StepCfg<BN>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  DSVC_CUDA(cudaOccupancyMaxActiveClusters(&n, tc_step_kernel<BN>, &cfg));
  *out = n;
  return DSVC_OK;
}

template <int BN>
int tc_step_launch(const StepPlan& plan, int pairs, cudaStream_t s) {
  DSVC_TRY((ensure_dyn_smem<tc_step_kernel<BN>>(StepCfg<BN>::SMEM)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(STEP_THREADS);
  cfg.dynamicSmemBytes = StepCfg<BN>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  DSVC_CUDA(cudaLaunchKernelEx(&cfg, tc_step_kernel<BN>, plan));
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
