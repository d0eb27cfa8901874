// fp32 FFMA implicit-GEMM for 1-D convolutions over channels-last activations.
//
//   D[b][op][n] = sum_{j < taps} sum_{ci < Cin} act(A[b][p*in_stride + in_off + j*tap_step][ci]) * W[j][n][ci]
//
// rows (p) are time positions, columns (n) output channels, K = taps x Cin.  Out-of-range input
// positions read as zero (the conv's zero padding).  One kernel covers Conv1d (any kernel size /
// dilation / stride), 1x1 projections (taps = 1) and ConvTranspose1d (one phase per blockIdx.z
// slice: taps = K/u, tap_step = -1, out position op = p*u + r).  The epilogue is a functor so the
// same main loop serves the WaveNet layers (fp32 validation mode) and the whole vocoder.
//
// This is the DSVC_MATH_FP32 path: exact fp32 products and fp32 accumulation.  The tcgen05 path
// (tc_gemm.cuh) shares the epilogue functors.
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace dsvc {

struct ConvGemmParams {
  const float* A;   // [B][Lin][Cin] channels-last
  const float* W;   // [nphase][taps][Cout][Cin]
  int B, Lin, Cin, Cout, taps;
  int rows;         // output rows per item (per phase)
  int in_stride, in_off, tap_step;
  int nphase;       // > 1: transposed conv with stride nphase; in_off is then (r + tpad) / nphase
  int tpad;
  float in_slope;   // leaky-relu slope applied to A on load (1.0f = identity)
  long long a_batch_stride;
};

template <int BM, int BN, int TM, int TN, bool VEC, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_gemm_f32_kernel(const ConvGemmParams p, const typename Epi::Params ep) {
  constexpr int BK = 16;
  constexpr int NTX = BN / TN;
  constexpr int NTY = BM / TM;
  constexpr int NTH = NTX * NTY;
  static_assert(TN == 4 || TN == 8, "TN must be 4 or 8");
  static_assert(TM == 4 || TM == 8, "TM must be 4 or 8");
  static_assert(!Epi::kPair || TN == 8, "pair epilogues need the two-chunk column mapping");
  constexpr int A_LD = VEC ? (BM * 4 + NTH - 1) / NTH : (BM * BK + NTH - 1) / NTH;
  constexpr int B_LD = VEC ? (BN * 4 + NTH - 1) / NTH : (BN * BK + NTH - 1) / NTH;

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int p0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int b = blockIdx.z, r = 0;
  int in_off = p.in_off;
  const float* Wp = p.W;
  if (p.nphase > 1) {
    b = blockIdx.z / p.nphase;
    r = blockIdx.z % p.nphase;
    in_off = (r + p.tpad) / p.nphase;
    Wp += (size_t)r * p.taps * p.Cout * p.Cin;
  }
  const float* Ab = p.A + (size_t)b * p.a_batch_stride;
  const int Ktot = p.taps * p.Cin;
  const int nchunks = (Ktot + BK - 1) / BK;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra4[VEC ? A_LD : 1], rb4[VEC ? B_LD : 1];
  float ras[VEC ? 1 : A_LD], rbs[VEC ? 1 : B_LD];

  auto load_regs = [&](int c) {
    if constexpr (VEC) {
      const int k0 = c * BK;
      const int j = k0 / p.Cin;
      const int cib = k0 - j * p.Cin;
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * NTH;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < BM * 4) {
          const int row = idx >> 2, q = idx & 3;
          const int pr = p0 + row;
          const int pos = pr * p.in_stride + in_off + j * p.tap_step;
          if (pr < p.rows && pos >= 0 && pos < p.Lin) {
            v = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)pos * p.Cin + cib + q * 4));
            if (p.in_slope != 1.0f) {
              v.x = lrelu_(v.x, p.in_slope); v.y = lrelu_(v.y, p.in_slope);
              v.z = lrelu_(v.z, p.in_slope); v.w = lrelu_(v.w, p.in_slope);
            }
          }
        }
        ra4[i] = v;
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * NTH;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < BN * 4) {
          const int n = n0 + (idx >> 2), q = idx & 3;
          if (n < p.Cout)
            v = __ldg(reinterpret_cast<const float4*>(Wp + ((size_t)j * p.Cout + n) * p.Cin + cib + q * 4));
        }
        rb4[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * NTH;
        float v = 0.f;
        if (idx < BM * BK) {
          const int row = idx / BK, kq = idx % BK;
          const int kk = c * BK + kq;
          const int pr = p0 + row;
          if (kk < Ktot && pr < p.rows) {
            const int j = kk / p.Cin, ci = kk - j * p.Cin;
            const int pos = pr * p.in_stride + in_off + j * p.tap_step;
            if (pos >= 0 && pos < p.Lin) {
              v = __ldg(Ab + (size_t)pos * p.Cin + ci);
              if (p.in_slope != 1.0f) v = lrelu_(v, p.in_slope);
            }
          }
        }
        ras[i] = v;
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * NTH;
        float v = 0.f;
        if (idx < BN * BK) {
          const int nn = idx / BK, kq = idx % BK;
          const int kk = c * BK + kq;
          const int n = n0 + nn;
          if (kk < Ktot && n < p.Cout) {
            const int j = kk / p.Cin, ci = kk - j * p.Cin;
            v = __ldg(Wp + ((size_t)j * p.Cout + n) * p.Cin + ci);
          }
        }
        rbs[i] = v;
      }
    }
  };

  auto store_smem = [&](int buf) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * NTH;
        if (idx < BM * 4) {
          const int row = idx >> 2, q = idx & 3;
          As[buf][q * 4 + 0][row] = ra4[i].x; As[buf][q * 4 + 1][row] = ra4[i].y;
          As[buf][q * 4 + 2][row] = ra4[i].z; As[buf][q * 4 + 3][row] = ra4[i].w;
        }
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * NTH;
        if (idx < BN * 4) {
          const int nn = idx >> 2, q = idx & 3;
          Bs[buf][q * 4 + 0][nn] = rb4[i].x; Bs[buf][q * 4 + 1][nn] = rb4[i].y;
          Bs[buf][q * 4 + 2][nn] = rb4[i].z; Bs[buf][q * 4 + 3][nn] = rb4[i].w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * NTH;
        if (idx < BM * BK) As[buf][idx % BK][idx / BK] = ras[i];
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        const int idx = tid + i * NTH;
        if (idx < BN * BK) Bs[buf][idx % BK][idx / BK] = rbs[i];
      }
    }
  };

  load_regs(0);
  store_smem(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) load_regs(c + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
      {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        bb[0] = v.x; bb[1] = v.y; bb[2] = v.z; bb[3] = v.w;
      }
      if constexpr (TN == 8) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + BN / 2]);
        bb[4] = v.x; bb[5] = v.y; bb[6] = v.z; bb[7] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (c + 1 < nchunks) store_smem(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: each thread owns TM rows x (one or two) 4-column chunks ----
  if constexpr (Epi::kPair) {
    const int c0 = blockIdx.y * (BN / 2) + tx * 4;   // pair-channel index
    const EpiCol cc = Epi::col(ep, c0);
    EpiPre pre[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int pr = p0 + ty * TM + i;
      if (pr < p.rows) pre[i] = Epi::pre(ep, b, pr, c0);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int pr = p0 + ty * TM + i;
      if (pr >= p.rows) continue;
      float g[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      float f[4] = {acc[i][4], acc[i][5], acc[i][6], acc[i][7]};
      Epi::apply_pair(ep, b, pr, c0, g, f, cc, pre[i]);
    }
  } else {
#pragma unroll
    for (int h2 = 0; h2 < TN / 4; ++h2) {
      const int n = n0 + tx * 4 + h2 * (BN / 2);
      if (n >= p.Cout) continue;
      const EpiCol cc = Epi::col(ep, n);
      EpiPre pre[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int pr = p0 + ty * TM + i;
        if (pr < p.rows) pre[i] = Epi::pre(ep, b, (p.nphase > 1) ? pr * p.nphase + r : pr, n);
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int pr = p0 + ty * TM + i;
        if (pr >= p.rows) continue;
        const int op = (p.nphase > 1) ? pr * p.nphase + r : pr;
        float v[4] = {acc[i][4 * h2], acc[i][4 * h2 + 1], acc[i][4 * h2 + 2], acc[i][4 * h2 + 3]};
        Epi::apply(ep, b, op, n, v, cc, pre[i]);
      }
    }
  }
}

// Host-side launcher.  tile: 0 = 128x128, 1 = 64x128, 2 = 128x64, 3 = 256x32, 4 = 256x16
template <int BM, int BN, int TM, int TN, class Epi>
int launch_conv_gemm_tile(const ConvGemmParams& p, const typename Epi::Params& ep, cudaStream_t s) {
  dim3 grid(ceil_div(p.rows, BM), ceil_div(p.Cout, BN), p.B * (p.nphase > 1 ? p.nphase : 1));
  dim3 block((BM / TM) * (BN / TN));
  const bool vec = (p.Cin % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);
  if (vec)
    conv_gemm_f32_kernel<BM, BN, TM, TN, true, Epi><<<grid, block, 0, s>>>(p, ep);
  else
    conv_gemm_f32_kernel<BM, BN, TM, TN, false, Epi><<<grid, block, 0, s>>>(p, ep);
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // namespace dsvc
