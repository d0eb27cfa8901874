// libdsvc: PitchExtractor (mel -> f0) of the 24 kHz models -- SURVEY.md section 8f row 4.
//
// Reference: modules/fastspeech/pe.py:120-149 (PitchExtractor.forward) =
//   Prenet (pe.py:8-44: 3 x [Conv1d k5 -> ReLU -> BatchNorm1d(eval)] * nonpadding, Linear, * nonpadding)
//   ConvStacks (pe.py:83-117: Linear, n x [x + ReLU(GroupNorm(Conv1d k5(x)))], Linear)
//   PitchPredictor (modules/fastspeech/tts_modules.py:192-235: + alpha * sinusoidal position embedding,
//                   5 x [pad, Conv1d k5 -> ReLU -> LayerNorm(channels, eps 1e-12)], Linear -> 2)
//   denorm_f0 (utils/pitch_utils.py:63-76).
// Called from Svc.infer when use_pe (infer_tools/infer_tool.py:164-165; infer.py:20 enables it for 24 kHz only).
//
// ~7 MFLOP per frame in fp32: it shares the FFMA implicit-GEMM of the vocoder (simt_gemm.cuh) with a frame-wise
// epilogue (bias, ReLU, folded BatchNorm, padding mask); the three normalisations that need a reduction the GEMM
// tile does not own (GroupNorm over time, LayerNorm over channels, the position scan) are small row kernels.
#include "simt_gemm.cuh"

namespace dsvc {

// ---- GEMM epilogue: v = acc + bias; relu; v = v * scale + shift (folded BatchNorm); v *= mask[b][t] ----------------
struct EpiFrame {
  static constexpr bool kPair = false;
  struct Params {
    const float* bias;    // [Cout]
    const float* scale;   // [Cout] or null
    const float* shift;   // [Cout] (with scale)
    const float* mask;    // [B][Lout] or null
    float* out;           // [B][Lout][Cout]
    int Lout, Cout, relu;
  };
  __device__ static __forceinline__ EpiCol col(const Params& e, int n) {
    EpiCol c;
    c.bias = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    c.d = make_float4(0.f, 0.f, 0.f, 0.f);
    return c;
  }
  __device__ static __forceinline__ void l2_prefetch(const Params&, int, int, int) {}
  __device__ static __forceinline__ EpiPre pre(const Params& e, int b, int op, int n) {
    EpiPre r{};
    if (e.scale) {
      r.a = __ldg(reinterpret_cast<const float4*>(e.scale + n));
      r.b = __ldg(reinterpret_cast<const float4*>(e.shift + n));
    }
    return r;
  }
  __device__ static __forceinline__ void apply(const Params& e, int b, int op, int n, const float (&a)[4], const EpiCol& c,
                                               const EpiPre& r) {
    float v[4] = {a[0] + c.bias.x, a[1] + c.bias.y, a[2] + c.bias.z, a[3] + c.bias.w};
    if (e.relu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    if (e.scale) {
      v[0] = add_rn(mul_rn(v[0], r.a.x), r.b.x); v[1] = add_rn(mul_rn(v[1], r.a.y), r.b.y);
      v[2] = add_rn(mul_rn(v[2], r.a.z), r.b.z); v[3] = add_rn(mul_rn(v[3], r.a.w), r.b.w);
    }
    if (e.mask) {
      const float m = e.mask[(size_t)b * e.Lout + op];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] *= m;
    }
    *reinterpret_cast<float4*>(e.out + ((size_t)b * e.Lout + op) * e.Cout + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// nonpadding[b][t] = (abs(mel[b][t]).sum() != 0)  (pe.py:31-32, :142).  One warp per frame.
__global__ void frame_mask_kernel(const float* __restrict__ mel, int rows, int M, float* __restrict__ mask) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.0f;
  for (int c = lane; c < M; c += 32) s += fabsf(mel[(size_t)row * M + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) mask[row] = (s == 0.0f) ? 0.0f : 1.0f;
}

// GroupNorm statistics over (C/G channels x T frames) per (item, group): stats[b][g] = (scale-free mean, rstd)
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ x, int T, int C, int G, float eps, float2* __restrict__ stats) {
  const int b = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
  const float* xb = x + (size_t)b * T * C + (size_t)g * cg;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < T * cg; i += blockDim.x) {
    const float v = xb[(size_t)(i / cg) * C + (i % cg)];
    s += v;
    ss += (double)v * v;
  }
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; q += sh[1][w]; }
    const double n = (double)T * cg, mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[blockIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// x = x + relu((c - mean) * rstd * gamma + beta)   (pe.py:64-76, :107-108)
__global__ void groupnorm_relu_residual_kernel(float* __restrict__ x, const float* __restrict__ c, const float2* __restrict__ stats,
                                               const float* __restrict__ gamma, const float* __restrict__ beta, int T, int C,
                                               int G, long long n4) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const long long e = i4 * 4;
  const int ch = (int)(e % C);
  const int b = (int)(e / ((long long)T * C));
  const float2 st = stats[b * G + ch / (C / G)];   // 4 consecutive channels share a group (C/G % 4 == 0)
  const float4 cv = reinterpret_cast<const float4*>(c)[i4];
  float4 xv = reinterpret_cast<float4*>(x)[i4];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch)), be = __ldg(reinterpret_cast<const float4*>(beta + ch));
  auto f = [&](float cc, float g, float bb) {
    return fmaxf(add_rn(mul_rn(mul_rn(sub_rn(cc, st.x), st.y), g), bb), 0.0f);
  };
  xv.x = add_rn(xv.x, f(cv.x, ga.x, be.x)); xv.y = add_rn(xv.y, f(cv.y, ga.y, be.y));
  xv.z = add_rn(xv.z, f(cv.z, ga.z, be.z)); xv.w = add_rn(xv.w, f(cv.w, ga.w, be.w));
  reinterpret_cast<float4*>(x)[i4] = xv;
}

// LayerNorm over channels, one warp per frame (tts_modules.py:37-56: eps 1e-12)
__global__ void layernorm_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int rows, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = in + (size_t)row * C;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += x[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.0f;
  for (int c = lane; c < C; c += 32) { const float d = x[c] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)C + eps);
  for (int c = lane; c < C; c += 32)
    out[(size_t)row * C + c] = add_rn(mul_rn(mul_rn(sub_rn(x[c], mean), rstd), gamma[c]), beta[c]);
}

// positions = cumsum(x[..., 0] != 0) * (x[..., 0] != 0)  (utils/__init__.py:145-157 with padding_idx 0); one warp per item
__global__ void positions_kernel(const float* __restrict__ x, int T, int C, int* __restrict__ pos) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int base = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool f = t < T && x[((size_t)b * T + t) * C] != 0.0f;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (t < T) pos[(size_t)b * T + t] = f ? base + __popc(bal & (0xffffffffu >> (31 - lane))) : 0;
    base += __popc(bal);
  }
}

// x += alpha * table[pos]  (tts_modules.py:228-229)
__global__ void add_positions_kernel(float* __restrict__ x, const int* __restrict__ pos, const float* __restrict__ table,
                                     const float* __restrict__ alpha, int C, long long n4) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const long long e = i4 * 4;
  const int ch = (int)(e % C);
  const long long row = e / C;
  const float a = alpha[0];
  const float4 tv = __ldg(reinterpret_cast<const float4*>(table + (size_t)pos[row] * C + ch));
  float4 xv = reinterpret_cast<float4*>(x)[i4];
  xv.x = add_rn(xv.x, mul_rn(a, tv.x)); xv.y = add_rn(xv.y, mul_rn(a, tv.y));
  xv.z = add_rn(xv.z, mul_rn(a, tv.z)); xv.w = add_rn(xv.w, mul_rn(a, tv.w));
  reinterpret_cast<float4*>(x)[i4] = xv;
}

// pitch_pred = linear(x) [odim <= 4]; f0_denorm = denorm_f0(pred[..., 0], uv = pred[..., 1] > 0, padding)  (pe.py:140-148)
__global__ void pitch_head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ mask, int rows, int C, int odim, int norm_mode, float f0_mean,
                                  float f0_std, int apply_uv, float* __restrict__ pred, float* __restrict__ f0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < C; c += 32) {
    const float v = x[(size_t)row * C + c];
    for (int o = 0; o < odim; ++o) acc[o] = fmaf(v, w[(size_t)o * C + c], acc[o]);
  }
  for (int o = 0; o < odim; ++o) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    acc[o] += bias[o];
  }
  if (lane == 0) {
    for (int o = 0; o < odim; ++o) pred[(size_t)row * odim + o] = acc[o];
    float v = acc[0];
    if (norm_mode == 2) v = add_rn(mul_rn(v, f0_std), f0_mean);     // 'standard'
    if (norm_mode == 1) v = powf(2.0f, v);                          // 'log'
    if (apply_uv && odim > 1 && acc[1] > 0.0f) v = 0.0f;
    if (mask[row] == 0.0f) v = 0.0f;
    f0[row] = v;
  }
}

}  // namespace dsvc

using namespace dsvc;

struct dsvc_pe {
  dsvc_pe_config cfg;
  // [taps][Cout][Cin] conv weights / [Cout][Cin] linears, biases, norms
  std::vector<DevBuf> w, b, n1, n2;   // indexed by the layer list below
  DevBuf pos_table, alpha, head_w, head_b;
  // activations
  DevBuf mask, xa, xb, tmp, stats, pos;
  int cap_rows = 0;
};

namespace {

enum { L_PRE0 = 0 };   // layer order: prenet convs, prenet out, enc in, enc convs, enc out, predictor convs

int upload_conv(DevBuf& dst, const float* w, int Cout, int Cin, int K, cudaStream_t s) {
  // PyTorch [Cout][Cin][K] -> [K][Cout][Cin]
  std::vector<float> r((size_t)K * Cout * Cin);
  for (int n = 0; n < Cout; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < K; ++j) r[((size_t)j * Cout + n) * Cin + c] = w[((size_t)n * Cin + c) * K + j];
  DSVC_TRY(dst.upload(r.data(), r.size() * sizeof(float), s));
  DSVC_CUDA(cudaStreamSynchronize(s));   // r dies here
  return DSVC_OK;
}

int launch_frame_gemm(const ConvGemmParams& p, const EpiFrame::Params& e, cudaStream_t s) {
  const int ctas = ceil_div(p.rows, 64) * ceil_div(p.Cout, 64) * p.B;
  if (ctas >= 4 * 148) return launch_conv_gemm_tile<128, 64, 8, 4, EpiFrame>(p, e, s);
  return launch_conv_gemm_tile<64, 64, 4, 4, EpiFrame>(p, e, s);
}

ConvGemmParams conv_params(const float* A, const float* W, int B, int T, int Cin, int Cout, int K, int pad_left) {
  ConvGemmParams p{};
  p.A = A; p.W = W; p.B = B; p.Lin = T; p.Cin = Cin; p.Cout = Cout; p.taps = K; p.rows = T;
  p.in_stride = 1; p.in_off = -pad_left; p.tap_step = 1; p.nphase = 1; p.tpad = 0; p.in_slope = 1.0f;
  p.a_batch_stride = (long long)T * Cin;
  return p;
}

}  // namespace

extern "C" {

int dsvc_pe_create(dsvc_pe_t** out, const dsvc_pe_config* cfg, const dsvc_pe_weights* w, void* stream) {
  DSVC_TRY(require_device());
  DSVC_REQUIRE(out && cfg && w, "dsvc_pe_create: null argument");
  const dsvc_pe_config& c = *cfg;
  DSVC_REQUIRE(c.n_mel_bins > 0 && c.n_mel_bins % 16 == 0 && c.hidden_size > 0 && c.hidden_size % 64 == 0 && c.predictor_hidden % 64 == 0,
               "dsvc_pe_create: n_mel_bins=%d must be a multiple of 16, hidden_size=%d / predictor_hidden=%d multiples of 64",
               c.n_mel_bins, c.hidden_size, c.predictor_hidden);
  DSVC_REQUIRE(c.prenet_layers >= 1 && c.prenet_layers <= 8 && c.enc_layers >= 0 && c.enc_layers <= 8 && c.pred_layers >= 1 && c.pred_layers <= 8,
               "dsvc_pe_create: layer counts out of range");
  DSVC_REQUIRE(c.prenet_kernel % 2 == 1 && c.enc_kernel % 2 == 1 && c.pred_kernel >= 1, "dsvc_pe_create: kernel sizes");
  DSVC_REQUIRE(c.enc_layers == 0 || (c.gn_groups > 0 && c.hidden_size % c.gn_groups == 0 && (c.hidden_size / c.gn_groups) % 4 == 0),
               "dsvc_pe_create: gn_groups=%d", c.gn_groups);
  DSVC_REQUIRE(c.odim >= 1 && c.odim <= 4 && c.pos_rows >= 2, "dsvc_pe_create: odim=%d pos_rows=%d", c.odim, c.pos_rows);
  cudaStream_t s = (cudaStream_t)stream;
  dsvc_pe* h = new dsvc_pe();
  h->cfg = c;
  const int H = c.hidden_size, P = c.predictor_hidden;
  const int nl = c.prenet_layers + 1 + (c.enc_layers > 0 ? c.enc_layers + 2 : 0) + c.pred_layers;
  h->w = std::vector<DevBuf>(nl); h->b = std::vector<DevBuf>(nl); h->n1 = std::vector<DevBuf>(nl); h->n2 = std::vector<DevBuf>(nl);
  int li = 0, rc = DSVC_OK;
  auto fail = [&](int r) { delete h; return r; };
#define PE_TRY(x) do { rc = (x); if (rc != DSVC_OK) return fail(rc); } while (0)
  for (int l = 0; l < c.prenet_layers; ++l, ++li) {
    const int cin = l == 0 ? c.n_mel_bins : H;
    PE_TRY(upload_conv(h->w[li], w->prenet_conv_w[l], H, cin, c.prenet_kernel, s));
    PE_TRY(h->b[li].upload(w->prenet_conv_b[l], H * sizeof(float), s));
    // BatchNorm1d eval folded the way ATen's CPU kernel does: alpha = invstd * weight, beta = bias - mean * alpha
    std::vector<float> alpha(H), beta(H);
    for (int i = 0; i < H; ++i) {
      const float invstd = 1.0f / sqrtf(w->prenet_bn_var[l][i] + c.bn_eps);
      alpha[i] = invstd * w->prenet_bn_w[l][i];
      beta[i] = w->prenet_bn_b[l][i] - w->prenet_bn_mean[l][i] * alpha[i];
    }
    PE_TRY(h->n1[li].upload(alpha.data(), H * sizeof(float), s));
    PE_TRY(h->n2[li].upload(beta.data(), H * sizeof(float), s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(DSVC_ECUDA);
  }
  PE_TRY(upload_conv(h->w[li], w->prenet_out_w, H, H, 1, s));
  PE_TRY(h->b[li].upload(w->prenet_out_b, H * sizeof(float), s));
  ++li;
  if (c.enc_layers > 0) {
    PE_TRY(upload_conv(h->w[li], w->enc_in_w, H, H, 1, s));
    PE_TRY(h->b[li].upload(w->enc_in_b, H * sizeof(float), s));
    ++li;
    for (int l = 0; l < c.enc_layers; ++l, ++li) {
      PE_TRY(upload_conv(h->w[li], w->enc_conv_w[l], H, H, c.enc_kernel, s));
      PE_TRY(h->b[li].upload(w->enc_conv_b[l], H * sizeof(float), s));
      PE_TRY(h->n1[li].upload(w->enc_gn_w[l], H * sizeof(float), s));
      PE_TRY(h->n2[li].upload(w->enc_gn_b[l], H * sizeof(float), s));
    }
    PE_TRY(upload_conv(h->w[li], w->enc_out_w, H, H, 1, s));
    PE_TRY(h->b[li].upload(w->enc_out_b, H * sizeof(float), s));
    ++li;
  }
  for (int l = 0; l < c.pred_layers; ++l, ++li) {
    const int cin = l == 0 ? H : P;
    PE_TRY(upload_conv(h->w[li], w->pred_conv_w[l], P, cin, c.pred_kernel, s));
    PE_TRY(h->b[li].upload(w->pred_conv_b[l], P * sizeof(float), s));
    PE_TRY(h->n1[li].upload(w->pred_ln_w[l], P * sizeof(float), s));
    PE_TRY(h->n2[li].upload(w->pred_ln_b[l], P * sizeof(float), s));
  }
  PE_TRY(h->head_w.upload(w->pred_linear_w, (size_t)c.odim * P * sizeof(float), s));
  PE_TRY(h->head_b.upload(w->pred_linear_b, c.odim * sizeof(float), s));
  PE_TRY(h->pos_table.upload(w->pos_table, (size_t)c.pos_rows * H * sizeof(float), s));
  PE_TRY(h->alpha.upload(w->pos_embed_alpha, sizeof(float), s));
  if (cudaStreamSynchronize(s) != cudaSuccess) return fail(DSVC_ECUDA);
#undef PE_TRY
  *out = h;
  return DSVC_OK;
}

void dsvc_pe_destroy(dsvc_pe_t* h) { delete h; }

int dsvc_pe_forward(dsvc_pe_t* h, const float* mel, int32_t B, int32_t T, float* pitch_pred, float* f0_denorm, void* stream) {
  DSVC_TRY(require_device());
  DSVC_REQUIRE(h && mel && pitch_pred && f0_denorm, "dsvc_pe_forward: null argument");
  DSVC_REQUIRE(B >= 0 && T >= 0, "dsvc_pe_forward: B=%d T=%d", B, T);
  if (B == 0 || T == 0) return DSVC_OK;
  const dsvc_pe_config& c = h->cfg;
  DSVC_REQUIRE(T + 1 <= c.pos_rows, "dsvc_pe_forward: T=%d needs a position table of %d rows, the handle has %d", T, T + 1, c.pos_rows);
  cudaStream_t s = (cudaStream_t)stream;
  const int H = c.hidden_size, P = c.predictor_hidden, Cmax = H > P ? H : P;
  const int rows = B * T;
  DSVC_TRY(h->mask.reserve((size_t)rows * sizeof(float)));
  DSVC_TRY(h->xa.reserve((size_t)rows * Cmax * sizeof(float)));
  DSVC_TRY(h->xb.reserve((size_t)rows * Cmax * sizeof(float)));
  DSVC_TRY(h->tmp.reserve((size_t)rows * Cmax * sizeof(float)));
  DSVC_TRY(h->pos.reserve((size_t)rows * sizeof(int)));
  if (c.enc_layers > 0) DSVC_TRY(h->stats.reserve((size_t)B * c.gn_groups * sizeof(float2)));
  float* mask = h->mask.as<float>();
  float *x = h->xa.as<float>(), *y = h->xb.as<float>(), *tmp = h->tmp.as<float>();

  frame_mask_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(mel, rows, c.n_mel_bins, mask);
  DSVC_LAUNCH_CHECK();
  int li = 0;
  const float* in = mel;
  int cin = c.n_mel_bins;
  // Prenet (pe.py:23-44)
  for (int l = 0; l < c.prenet_layers; ++l, ++li) {
    EpiFrame::Params e{h->b[li].as<float>(), h->n1[li].as<float>(), h->n2[li].as<float>(), mask, y, T, H, 1};
    DSVC_TRY(launch_frame_gemm(conv_params(in, h->w[li].as<float>(), B, T, cin, H, c.prenet_kernel, c.prenet_kernel / 2), e, s));
    std::swap(x, y);
    in = x;
    cin = H;
  }
  {
    EpiFrame::Params e{h->b[li].as<float>(), nullptr, nullptr, mask, y, T, H, 0};
    DSVC_TRY(launch_frame_gemm(conv_params(x, h->w[li].as<float>(), B, T, H, H, 1, 0), e, s));
    std::swap(x, y);
    ++li;
  }
  // ConvStacks (pe.py:100-117)
  if (c.enc_layers > 0) {
    {
      EpiFrame::Params e{h->b[li].as<float>(), nullptr, nullptr, nullptr, y, T, H, 0};
      DSVC_TRY(launch_frame_gemm(conv_params(x, h->w[li].as<float>(), B, T, H, H, 1, 0), e, s));
      std::swap(x, y);
      ++li;
    }
    for (int l = 0; l < c.enc_layers; ++l, ++li) {
      EpiFrame::Params e{h->b[li].as<float>(), nullptr, nullptr, nullptr, tmp, T, H, 0};
      DSVC_TRY(launch_frame_gemm(conv_params(x, h->w[li].as<float>(), B, T, H, H, c.enc_kernel, c.enc_kernel / 2), e, s));
      groupnorm_stats_kernel<<<B * c.gn_groups, 256, 0, s>>>(tmp, T, H, c.gn_groups, c.gn_eps, h->stats.as<float2>());
      DSVC_LAUNCH_CHECK();
      const long long n4 = (long long)rows * H / 4;
      groupnorm_relu_residual_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(x, tmp, h->stats.as<float2>(), h->n1[li].as<float>(),
                                                                                  h->n2[li].as<float>(), T, H, c.gn_groups, n4);
      DSVC_LAUNCH_CHECK();
    }
    {
      EpiFrame::Params e{h->b[li].as<float>(), nullptr, nullptr, nullptr, y, T, H, 0};
      DSVC_TRY(launch_frame_gemm(conv_params(x, h->w[li].as<float>(), B, T, H, H, 1, 0), e, s));
      std::swap(x, y);
      ++li;
    }
  }
  // PitchPredictor (tts_modules.py:222-235)
  positions_kernel<<<B, 32, 0, s>>>(x, T, H, h->pos.as<int>());
  DSVC_LAUNCH_CHECK();
  {
    const long long n4 = (long long)rows * H / 4;
    add_positions_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(x, h->pos.as<int>(), h->pos_table.as<float>(), h->alpha.as<float>(), H, n4);
    DSVC_LAUNCH_CHECK();
  }
  cin = H;
  for (int l = 0; l < c.pred_layers; ++l, ++li) {
    const int pad_left = c.pad_same ? (c.pred_kernel - 1) / 2 : c.pred_kernel - 1;
    EpiFrame::Params e{h->b[li].as<float>(), nullptr, nullptr, nullptr, tmp, T, P, 1};
    DSVC_TRY(launch_frame_gemm(conv_params(x, h->w[li].as<float>(), B, T, cin, P, c.pred_kernel, pad_left), e, s));
    layernorm_rows_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(tmp, x, h->n1[li].as<float>(), h->n2[li].as<float>(), rows, P, c.ln_eps);
    DSVC_LAUNCH_CHECK();
    cin = P;
  }
  pitch_head_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, h->head_w.as<float>(), h->head_b.as<float>(), mask, rows, P, c.odim,
                                                       c.pitch_norm, c.f0_mean, c.f0_std, c.apply_uv, pitch_pred, f0_denorm);
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // extern "C"
