// libdsvc: the data formats either side of the hot path.
//
//   dsvc_mel_analysis   waveform -> log10-mel: reflect pad, framed Hann STFT, magnitude, mel filterbank, log-clamp,
//                       rescale.  Follows STFT.get_mel (modules/nsf_hifigan/nvSTFT.py:72-104) and the 0.434294
//                       rescale + transpose of NsfHifiGAN.wav2spec (network/vocoders/nsf_hifigan.py:76-92).
//   dsvc_compact_frames the mel/f0 glue of Svc.after_infer (infer_tools/infer_tool.py:172-200): drop all-zero
//                       (padding) frames, clip the mel to [mel_vmin, mel_vmax], keep f0 on the same frames -- on the
//                       device, so the denoised mel never leaves HBM between the sampler and the vocoder.
//
// Both are HBM/latency-trivial next to the sampler (a 10 s clip is 862 frames); they exist to remove the
// D2H -> numpy -> H2D hops, not to win FLOPs.  The DFT runs in fp64 (2048-point radix-2 in shared memory, twiddles
// from sincospi) so that the spectrum is correctly rounded fp32: the reference's own fp32 FFT (pocketfft / cuFFT)
// differs from the exact transform by more than this kernel does.
#include "common.cuh"

namespace dsvc {

// ---- dsvc_mel_analysis -----------------------------------------------------------------------------------------

// One CTA per frame.  smem: double2 buf[n_fft] | double2 tw[n_fft/2] | float mag[n_fft/2+1]
__global__ void __launch_bounds__(256)
stft_mel_kernel(const float* __restrict__ wav, long long n_samples, int n_fft, int log2n, int hop, int pad,
                const float* __restrict__ window, const float* __restrict__ basis, const int* __restrict__ band_lo,
                const int* __restrict__ band_hi, int n_mels, float clip_val, float out_scale,
                float* __restrict__ out /* [T][n_mels] */) {
  extern __shared__ __align__(16) unsigned char smem_mel[];
  double2* buf = reinterpret_cast<double2*>(smem_mel);
  double2* tw = buf + n_fft;
  float* mag = reinterpret_cast<float*>(tw + n_fft / 2);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long long frame = blockIdx.x;
  const int n_bins = n_fft / 2 + 1;

  // twiddles exp(-2 pi i k / N), k < N/2
  for (int k = tid; k < n_fft / 2; k += nthr) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)n_fft, &s, &c);
    tw[k] = make_double2(c, s);
  }
  // windowed frame of the reflect-padded signal (nvSTFT.py:91-92, torch.stft center=False), bit-reversed order.
  // The window multiply is the reference's fp32 product (torch.stft multiplies fp32 frames by the fp32 window).
  for (int n = tid; n < n_fft; n += nthr) {
    long long i = frame * hop + n - pad;
    if (i < 0) i = -i;
    if (i >= n_samples) i = 2 * (n_samples - 1) - i;
    const float v = mul_rn(wav[i], window[n]);
    const int r = (int)(__brev((unsigned)n) >> (32 - log2n));
    buf[r] = make_double2((double)v, 0.0);
  }
  __syncthreads();
  // in-place radix-2 decimation-in-time
  for (int s = 1; s <= log2n; ++s) {
    const int half = 1 << (s - 1);
    const int tstride = n_fft >> s;
    for (int j = tid; j < n_fft / 2; j += nthr) {
      const int grp = j >> (s - 1), k = j & (half - 1);
      const int i0 = (grp << s) + k, i1 = i0 + half;
      const double2 w = tw[k * tstride];
      const double2 a = buf[i0], b = buf[i1];
      const double tr = b.x * w.x - b.y * w.y, ti = b.x * w.y + b.y * w.x;
      buf[i0] = make_double2(a.x + tr, a.y + ti);
      buf[i1] = make_double2(a.x - tr, a.y - ti);
    }
    __syncthreads();
  }
  // magnitude: sqrt(re^2 + im^2 + 1e-9) on the fp32 spectrum, op order of nvSTFT.py:97
  for (int k = tid; k < n_bins; k += nthr) {
    const float re = (float)buf[k].x, im = (float)buf[k].y;
    mag[k] = sqrtf(add_rn(add_rn(mul_rn(re, re), mul_rn(im, im)), 1e-9f));
  }
  __syncthreads();
  // mel filterbank row dot (nvSTFT.py:99) over the row's non-zero band, log-clamp (:101, :56), rescale
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int m = warp; m < n_mels; m += nwarps) {
    const int lo = band_lo ? band_lo[m] : 0, hi = band_hi ? band_hi[m] : n_bins;
    const float* row = basis + (size_t)m * n_bins;
    float acc = 0.0f;
    for (int k = lo + lane; k < hi; k += 32) acc = fmaf(row[k], mag[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[frame * n_mels + m] = mul_rn(out_scale, logf(fmaxf(acc, clip_val)));
  }
}

// ---- dsvc_compact_frames ---------------------------------------------------------------------------------------

// One CTA (1024 threads = 32 warps); rows are visited in chunks of 1024 with a running output offset.
__global__ void __launch_bounds__(1024)
compact_frames_kernel(const float* __restrict__ mel, const float* __restrict__ f0, int T, int M, float vmin, float vmax,
                      float* __restrict__ mel_out, float* __restrict__ f0_out, int* __restrict__ n_kept) {
  __shared__ int flag[1024];
  __shared__ int pos[1024];
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int r0 = 0; r0 < T; r0 += 1024) {
    // keep = abs(mel).sum(-1) > 0 (infer_tool.py:183-184): each warp sums 32 rows
    for (int i = 0; i < 32; ++i) {
      const int r = r0 + warp * 32 + i;
      float s = 0.0f;
      if (r < T)
        for (int c = lane; c < M; c += 32) s += fabsf(mel[(size_t)r * M + c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) flag[warp * 32 + i] = (r < T && s > 0.0f) ? 1 : 0;
    }
    __syncthreads();
    // exclusive scan of the 1024 flags
    const int f = flag[tid];
    int incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += v;
      }
      warp_tot[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    pos[tid] = base + warp_tot[warp] + incl - f;
    __syncthreads();
    const int chunk_total = pos[1023] + flag[1023] - base;
    // scatter: clip(mel, vmin, vmax) (:186), f0 on the same rows (:193)
    for (int i = 0; i < 32; ++i) {
      const int q = warp * 32 + i, r = r0 + q;
      if (r < T && flag[q]) {
        const int d = pos[q];
        for (int c = lane; c < M; c += 32) mel_out[(size_t)d * M + c] = fminf(fmaxf(mel[(size_t)r * M + c], vmin), vmax);
        if (lane == 0 && f0 != nullptr) f0_out[d] = f0[r];
      }
    }
    __syncthreads();
    if (tid == 0) base += chunk_total;
    __syncthreads();
  }
  if (tid == 0) *n_kept = base;
}

}  // namespace dsvc

using namespace dsvc;

extern "C" {

int64_t dsvc_mel_frames(const dsvc_mel_config* cfg, int64_t n_samples) {
  if (!cfg || cfg->n_fft <= 0 || cfg->hop_size <= 0 || cfg->hop_size > cfg->n_fft) return -1;
  const int64_t pad = (cfg->n_fft - cfg->hop_size) / 2;
  if (n_samples <= pad) return -1;                       // reflect padding needs pad < n_samples
  const int64_t padded = n_samples + 2 * pad;
  if (padded < cfg->n_fft) return 0;
  return 1 + (padded - cfg->n_fft) / cfg->hop_size;
}

int dsvc_mel_analysis(const dsvc_mel_config* cfg, const float* wav, int64_t n_samples, const float* window,
                      const float* mel_basis, const int32_t* band_lo, const int32_t* band_hi, float* mel_out,
                      void* stream) {
  DSVC_TRY(require_device());
  DSVC_REQUIRE(cfg && wav && window && mel_basis && mel_out, "dsvc_mel_analysis: null argument");
  const int n = cfg->n_fft;
  DSVC_REQUIRE(n >= 64 && n <= 4096 && (n & (n - 1)) == 0, "dsvc_mel_analysis: n_fft=%d must be a power of two in [64, 4096]", n);
  DSVC_REQUIRE(cfg->n_mels > 0, "dsvc_mel_analysis: n_mels=%d", cfg->n_mels);
  DSVC_REQUIRE((band_lo == nullptr) == (band_hi == nullptr), "dsvc_mel_analysis: band_lo and band_hi go together");
  const int64_t T = dsvc_mel_frames(cfg, n_samples);
  DSVC_REQUIRE(T >= 0, "dsvc_mel_analysis: %lld samples cannot be reflect-padded by (n_fft - hop)/2 = %d", (long long)n_samples,
               (cfg->n_fft - cfg->hop_size) / 2);
  if (T == 0) return DSVC_OK;
  DSVC_REQUIRE(T < (1ll << 31), "dsvc_mel_analysis: too many frames");
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  const size_t smem = (size_t)n * sizeof(double2) + (size_t)(n / 2) * sizeof(double2) + (size_t)(n / 2 + 1) * sizeof(float);
  DSVC_TRY((ensure_dyn_smem<stft_mel_kernel>((int)smem)));
  stft_mel_kernel<<<(unsigned)T, 256, smem, (cudaStream_t)stream>>>(wav, (long long)n_samples, n, log2n, cfg->hop_size,
                                                                   (cfg->n_fft - cfg->hop_size) / 2, window, mel_basis, band_lo,
                                                                   band_hi, cfg->n_mels, cfg->clip_val, cfg->out_scale, mel_out);
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

int dsvc_compact_frames(const float* mel, const float* f0, int32_t T, int32_t M, float vmin, float vmax, float* mel_out,
                        float* f0_out, int32_t* n_kept, void* stream) {
  DSVC_TRY(require_device());
  DSVC_REQUIRE(mel && mel_out && n_kept, "dsvc_compact_frames: null argument");
  DSVC_REQUIRE((f0 == nullptr) == (f0_out == nullptr), "dsvc_compact_frames: f0 and f0_out go together");
  DSVC_REQUIRE(T >= 0 && M > 0, "dsvc_compact_frames: T=%d M=%d", T, M);
  compact_frames_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mel, f0, T, M, vmin, vmax, mel_out, f0_out, n_kept);
  DSVC_LAUNCH_CHECK();
  return DSVC_OK;
}

}  // extern "C"
