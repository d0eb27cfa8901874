// One WaveNet residual layer as ONE kernel (reference network/diff/net.py:66-84):
//
//   phase A  dilated conv k=3 (+ hoisted conditioner) -> sigmoid(gate) * tanh(filter) -> Z     (EpiGate)
//   ---- barrier.cluster (release / acquire) ----
//   phase B  1x1 output projection of Z -> residual /sqrt(2), skip accumulation, next layer's x + d   (EpiOutProj)
//
// Why a cluster is enough: the output projection is a 1x1 contraction, so a 128-frame tile of phase B reads the
// gated activation Z of the SAME 128 frames only -- all C channels of it, i.e. exactly what the 2C/64 CTAs that own
// this frame tile's 64-wide channel tiles produced in phase A.  Those CTAs form one thread-block cluster along the
// channel-tile axis (12 CTAs for C = 384: a non-portable cluster size, one cluster per GPC); Z travels through L2
// (generic-proxy stores -> fence.proxy.async -> cluster barrier -> TMA loads), nothing goes through DSMEM.
// Both phases are the main loop and the epilogues of tc_gemm.cuh (BN = 64, 2 MMAs per K-step in 3-pass mode,
// smem-staged 16-warp epilogue): barriers, TMEM and the smem ring are set up once and the pipeline iteration
// counter simply runs on through the second phase.  The arithmetic (MMA order, epilogue functors) is the one of the
// two separate kernels, so results are bit-identical to them.
//
// The write-after-read hazard fusion would otherwise create -- phase B of a fast cluster overwriting the conv-input
// plane Y while a neighbouring cluster's phase A still reads its +-dil halo rows from it -- is removed by ping-ponging
// Y between two planes by layer parity (diffnet.cu: `pingpong`).
//
// Measured on B200 (profiles/r1f_fused_layer_timeline.txt, r1f_fused_layer_ab.txt; one 862-frame clip): bit-identical;
// the in-kernel hand-over (last Z store -> first out-proj operands in smem) takes 4370 cycles = proxy fence 1300-1700 +
// cluster barrier 1400 + TMA from L2 1300, against ~6500 for the PDL kernel boundary it replaces; but the 12-CTA
// clusters fill 7 of the 8 GPCs, so the next layer's kernel cannot pre-launch next to this one.  Net: 17.9 vs 18.4 us
// per layer back to back, 381.4 vs 384.1 us per DDPM step -- within noise, hence opt-in (DSVC_FUSED_LAYER=1|2);
// DESIGN.md section 3.1c.
#pragma once
#include "tc_gemm.cuh"

namespace dsvc {

constexpr int LY_BN = 64;

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tmYh, const __grid_constant__ CUtensorMap tmYl,    // conv input planes
                const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,    // conv weights, box {64, 32}
                const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,    // gated activation planes
                const __grid_constant__ CUtensorMap tmOh, const __grid_constant__ CUtensorMap tmOl,    // out-proj weights, box {64, 32}
                const __grid_constant__ CUtensorMap tmNh, const __grid_constant__ CUtensorMap tmNl,    // NEXT layer's conv weights (L2 prefetch)
                const EpiGate::Params eg, const EpiOutProj::Params eo, int T, int K, int N, int dil, int passes, int prefetch_next, int light_fence) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  constexpr int BN = LY_BN;
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
#ifdef DSVC_TIMELINE
  const long long tl0 = clock64();
#endif
  const int m0 = blockIdx.x * TC_BM, ny = blockIdx.y, b = blockIdx.z;
  const int kblocks = K / TC_BK;
  const int total_a = 3 * kblocks;           // phase A: three taps
  const int total_b = kblocks;               // phase B: 1x1
  const bool three = passes == 3;
  const uint32_t tx_bytes = three ? Cfg::STAGE : Cfg::STAGE / 2;

  if (warp == 0 && lane == 0) {
    if (prefetch_next) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmNh) : "memory");
      if (three) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmNl) : "memory");
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOh) : "memory");
    if (three) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOl) : "memory");
    }
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
  if (warp == 3) TL_MARK(0);   // setup done

  // ---- operand loads ---------------------------------------------------------------------------
  // phase A weights: pair tiles live in a 128-row super-tile packed [64 gate rows | 64 filter rows]; CTA ny takes gate
  // rows 128*(ny/2) + 32*(ny%2) and the filter rows 64 further on (same packing as tc_gemm_kernel<EpiGate, 64>)
  auto load_wa = [&](int it, int s) {
    const int tap = it / kblocks, kb = it - tap * kblocks;
    const int r0 = tap * N + (ny >> 1) * 128 + (ny & 1) * 32, r1 = r0 + 64;
    tma_load_2d(&tmWh, full_bar(s), tile_b(s, 0), kb * TC_BK, r0);
    tma_load_2d(&tmWh, full_bar(s), tile_b(s, 0) + 32u * 128u, kb * TC_BK, r1);
    if (three) {
      tma_load_2d(&tmWl, full_bar(s), tile_b(s, 1), kb * TC_BK, r0);
      tma_load_2d(&tmWl, full_bar(s), tile_b(s, 1) + 32u * 128u, kb * TC_BK, r1);
    }
  };
  auto load_ya = [&](int it, int s) {
    const int tap = it / kblocks, kb = it - tap * kblocks;
    const int frame = m0 + (tap - 1) * dil;               // out-of-range frames are zero-filled = the conv's padding
    tma_load_3d(&tmYh, full_bar(s), tile_a(s, 0), kb * TC_BK, frame, b);
    if (three) tma_load_3d(&tmYl, full_bar(s), tile_a(s, 1), kb * TC_BK, frame, b);
  };
  // phase B: plain 64-row weight tile (rows ny*64, ny*64+32), Z rows of this frame tile
  auto load_b = [&](int kb, int s) {
    const int row = ny * BN;
    tma_load_3d(&tmZh, full_bar(s), tile_a(s, 0), kb * TC_BK, m0, b);
    if (three) tma_load_3d(&tmZl, full_bar(s), tile_a(s, 1), kb * TC_BK, m0, b);
    tma_load_2d(&tmOh, full_bar(s), tile_b(s, 0), kb * TC_BK, row);
    tma_load_2d(&tmOh, full_bar(s), tile_b(s, 0) + 32u * 128u, kb * TC_BK, row + 32);
    if (three) {
      tma_load_2d(&tmOl, full_bar(s), tile_b(s, 1), kb * TC_BK, row);
      tma_load_2d(&tmOl, full_bar(s), tile_b(s, 1) + 32u * 128u, kb * TC_BK, row + 32);
    }
  };
  // the MMAs of one pipeline stage (identical to tc_gemm_kernel, BN = 64)
  const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
  const uint32_t idesc2 = umma_idesc_f16(TC_BM, 2 * BN);   // [wh ; wl] concatenated along N
  auto issue_stage = [&](int s, bool first) {
    const uint64_t ah = umma_desc_sw128(tile_a(s, 0)), al = umma_desc_sw128(tile_a(s, 1));
    const uint64_t bh = umma_desc_sw128(tile_b(s, 0));
#pragma unroll
    for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
      const uint64_t koff = (uint64_t)((k4 * 32) >> 4);
      const uint32_t acc = (!first || k4 > 0) ? 1u : 0u;
      if (three) {
        umma_f16(tmem_base, ah + koff, bh + koff, idesc2, acc);
        umma_f16(tmem_base, al + koff, bh + koff, idesc, 1u);
      } else {
        umma_f16(tmem_base, ah + koff, bh + koff, idesc, acc);
      }
    }
  };

  // =================================== phase A: dilated conv + gate ===================================
  if (warp == 0) {
    const int pre = total_a < STAGES ? total_a : STAGES;
    if (elect_one_sync()) {
      for (int it = 0; it < pre; ++it) {          // stages are initially free; weights do not depend on the previous kernel
        mbar_expect_tx(full_bar(it), tx_bytes);
        load_wa(it, it);
      }
    }
    __syncwarp();
    pdl_wait();                                   // the conv-input plane was written by the previous kernel
    if (elect_one_sync()) {
      for (int it = 0; it < pre; ++it) load_ya(it, it);
      // pull phase B's weight tiles towards L2 while phase A runs (they are read right after the cluster barrier)
      for (int kb = 0; kb < kblocks; ++kb) {
        tma_prefetch_2d(&tmOh, kb * TC_BK, ny * BN);
        tma_prefetch_2d(&tmOh, kb * TC_BK, ny * BN + 32);
        if (three) {
          tma_prefetch_2d(&tmOl, kb * TC_BK, ny * BN);
          tma_prefetch_2d(&tmOl, kb * TC_BK, ny * BN + 32);
        }
      }
    }
    __syncwarp();
    for (int it = pre; it < total_a; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(full_bar(s), tx_bytes);
        load_ya(it, s);
        load_wa(it, s);
      }
      __syncwarp();
    }
    // Optional (DSVC_FUSED_PREFETCH=1; measured: no effect on the step time): the next layer's kernel cannot pre-launch
    // next to this one (its 12-CTA clusters need the SMs this grid holds), so its weight tiles are not requested early by
    // its own prologue.  Pull the tiles CTA (.., ny, ..) of that kernel will read into L2 from here -- this warp is
    // idle until the accumulator is complete (issuing them after phase B's loads delayed the epilogue by ~700 cycles).
    if (prefetch_next && elect_one_sync()) {
      for (int it = 0; it < total_a; ++it) {
        const int tap = it / kblocks, kb = it - tap * kblocks;
        const int r0 = tap * N + (ny >> 1) * 128 + (ny & 1) * 32;
        tma_prefetch_2d(&tmNh, kb * TC_BK, r0);
        tma_prefetch_2d(&tmNh, kb * TC_BK, r0 + 64);
        if (three) {
          tma_prefetch_2d(&tmNl, kb * TC_BK, r0);
          tma_prefetch_2d(&tmNl, kb * TC_BK, r0 + 64);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    for (int it = 0; it < total_a; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      if (it == 0) TL_MARK(1);             // first operands of phase A landed
      tc_fence_after();
      if (elect_one_sync()) {
        issue_stage(s, it == 0);
        umma_commit(empty_bar(s));
        if (it == total_a - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
    TL_MARK(2);                            // phase A MMAs issued
  }
  pdl_wait();   // every warp: the epilogues read and overwrite tensors of the previous kernel
#ifdef DSVC_TIMELINE
  tc_epilogue<EpiGate, BN>(eg, smem_raw, smem_base, tmem_base, tmem_full_bar, 0u, T, N, m0, ny, b, warp, lane, three, tl0);
#else
  tc_epilogue<EpiGate, BN>(eg, smem_raw, smem_base, tmem_base, tmem_full_bar, 0u, T, N, m0, ny, b, warp, lane, three);
#endif

  // ============ Z of this frame tile is complete once every CTA of the cluster has passed here ============
  if (warp == 4) TL_MARK(6);                            // epilogue A done
  // light_fence (default): rely on the cluster barrier's own release / acquire for the visibility of the Z stores and
  // keep only the proxy fences; DSVC_FUSED_FENCE=1: a device-scope fence per thread first
  if (!light_fence) __threadfence();                    // this thread's Z stores are performed device-wide ...
  asm volatile("fence.proxy.async;" ::: "memory");      // ... and ordered against the peers' TMA (async-proxy) reads
  if (warp == 4) TL_MARK(7);                            // fences done
  tc_fence_before();                                    // TMEM reads of phase A precede phase B's MMAs
  __syncwarp();
  cluster_sync_all();                                   // also: every warp is done with the staging slab (smem ring)
  tc_fence_after();
  if (warp == 4) TL_MARK(8);                            // cluster barrier passed

  // =================================== phase B: output projection ===================================
  if (warp == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    for (int it = 0; it < total_b; ++it) {
      const int g = total_a + it;                 // the pipeline iteration counter runs on
      const int s = g % STAGES;
      const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);           // g >= STAGES always: phase A has at least STAGES iterations
      if (elect_one_sync()) {
        mbar_expect_tx(full_bar(s), tx_bytes);
        load_b(it, s);
      }
      __syncwarp();
    }
    __syncwarp();
  } else if (warp == 1) {
    for (int it = 0; it < total_b; ++it) {
      const int g = total_a + it;
      const int s = g % STAGES;
      const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      if (it == 0) TL_MARK(9);             // first operands of phase B landed
      tc_fence_after();
      if (elect_one_sync()) {
        issue_stage(s, it == 0);
        umma_commit(empty_bar(s));
        if (it == total_b - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
    TL_MARK(10);                           // phase B MMAs issued
  }
#ifdef DSVC_TIMELINE
  tc_epilogue<EpiOutProj, BN>(eo, smem_raw, smem_base, tmem_base, tmem_full_bar, 1u, T, N, m0, ny, b, warp, lane, three, tl0, 8);
  if (warp == 4) TL_MARK(14);              // epilogue B done
#else
  tc_epilogue<EpiOutProj, BN>(eo, smem_raw, smem_base, tmem_base, tmem_full_bar, 1u, T, N, m0, ny, b, warp, lane, three);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
#endif
}

// DSVC_FUSED_LAYER: unset / 0 = off, 1 = where the two separate kernels would run 64-wide tiles (one clip, real-time
// chunks: the launch-bound regime), 2 = always.  Read per call (handle creation decides the plane ping-pong, graph
// capture decides the kernels): tests switch it per handle, like DSVC_SPLITK.
inline int tc_layer_env() {
  const char* e = getenv("DSVC_FUSED_LAYER");
  return e ? atoi(e) : 0;
}

// shape rule only (the cluster must also be schedulable: tc_layer_launch reports that)
inline bool tc_layer_shape_ok(int B, int T, int C) {
  const int env = tc_layer_env();
  if (env <= 0) return false;
  const int N = 2 * C;
  if (C % TC_BK != 0 || N % 128 != 0 || N / LY_BN > 16 || 3 * (C / TC_BK) < TcCfg<LY_BN>::STAGES) return false;
  if (env >= 2) return true;
  return tc_pick_bn(B, T, N) == 64;
}

// Can a cluster of nt = 2C/64 CTAs of this kernel (193 KB of shared memory each) be co-scheduled on the current device?
// Sets the function attributes on first use; cached per cluster size.  Called from dsvc_diffnet_prepare (outside any
// stream capture).  *usable = 0: the caller keeps the two separate kernels.
inline int tc_layer_probe(int nt, int* usable) {   // *usable = max co-resident clusters (0: not schedulable)
  static int cache[64][17];                        // per device: 0 = not probed yet, -1 = not schedulable, else clusters
  *usable = 0;
  if (nt < 1 || nt > 16) return DSVC_OK;
  int dev = 0;
  DSVC_CUDA(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < 64;
  int v = tracked ? cache[dev][nt] : 0;
  if (v == 0) {
    auto kern = tc_layer_kernel;
    DSVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<LY_BN>::SMEM));
    if (nt > 8) DSVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(1, nt, 1);
    q.blockDim = dim3(TC_THREADS);
    q.dynamicSmemBytes = TcCfg<LY_BN>::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = nt;
    attr[0].val.clusterDim.z = 1;
    q.attrs = attr;
    q.numAttrs = 1;
    int clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&clusters, kern, &q);
    if (e != cudaSuccess) cudaGetLastError();
    v = (e == cudaSuccess && clusters >= 1) ? clusters : -1;
    if (tracked) cache[dev][nt] = v;
  }
  *usable = v > 0 ? v : 0;
  return DSVC_OK;
}

// one launch = one residual layer; the caller has checked tc_layer_shape_ok() and tc_layer_probe()
// `mnext`: the next layer's conv-weight maps for the L2 prefetch (null: last layer / prefetch off)
inline bool tc_layer_prefetch_next() {
  const char* e = getenv("DSVC_FUSED_PREFETCH");     // default off (no measured effect); 1 switches the next-layer weight prefetch on
  return e && e[0] == '1';
}
inline bool tc_layer_light_fence() {
  // default: proxy fences + the cluster barrier's release / acquire (bit-identical in every test, ~500 cycles shorter);
  // DSVC_FUSED_FENCE=1 adds a device-scope fence per thread in front
  const char* e = getenv("DSVC_FUSED_FENCE");
  return !(e && e[0] == '1');
}
inline int tc_layer_launch(const TcGemmMaps& md, const TcGemmMaps& mo, const TcGemmMaps* mnext, const EpiGate::Params& eg,
                           const EpiOutProj::Params& eo, int B, int T, int C, int dil, int passes, cudaStream_t s) {
  const int N = 2 * C, nt = N / LY_BN;
  DSVC_REQUIRE(nt >= 1 && nt <= 16 && N % 128 == 0 && C % TC_BK == 0, "tc_layer_launch: C=%d does not tile into a cluster", C);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div(T, TC_BM), nt, B);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<LY_BN>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 1;
  attr[na].val.clusterDim.y = nt;
  attr[na].val.clusterDim.z = 1;
  ++na;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const bool pf = mnext != nullptr && tc_layer_prefetch_next();
  const TcGemmMaps& mn = pf ? *mnext : md;
  DSVC_CUDA(cudaLaunchKernelEx(&cfg, tc_layer_kernel, md.a_hi, md.a_lo, md.b32_hi, md.b32_lo, mo.a_hi, mo.a_lo, mo.b32_hi, mo.b32_lo,
                               mn.b32_hi, mn.b32_lo, eg, eo, T, C, N, dil, passes, pf ? 1 : 0, tc_layer_light_fence() ? 1 : 0));
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
