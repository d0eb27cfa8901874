"""CPU oracle for the diffusion-SVC inference hot path.  TEST INFRASTRUCTURE ONLY.

A from-scratch, functional (state-dict in, tensors out) restatement in plain
PyTorch-on-CPU of what the reference computes on the hot path named by
BASELINE.json's north_star.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module;
the product (`diffsvc_b200`) never does and fails loudly without its CUDA
library.

Parity pinning: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference's own modules imported on CPU in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`, committed) and, when
/root/reference is present, re-checked live by
`tests/test_oracle_vs_reference.py`.

All file:line citations are into /root/reference (prophesier/diff-svc @ 76154f0).
The arithmetic itself (conv1d / conv_transpose1d / linear) is delegated to
torch.nn.functional on CPU fp32 (or fp64 with `dtype=torch.float64`), exactly
the third-party library the reference calls at those sites.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules/nsf_hifigan/models.py:12


# --------------------------------------------------------------------------------------
# DiffNet (network/diff/net.py)
# --------------------------------------------------------------------------------------

def sinusoidal_pos_emb(t, dim):
    """network/diff/net.py:37-44.  t: int64/float [B] -> [B, dim] (cat(sin, cos))."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -k)
    emb = t[:, None] * freqs[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def mish(x):
    """modules/commons/common_layers.py:485-487."""
    return x * torch.tanh(F.softplus(x))


def diffnet_dims(sd):
    """Infer (M, C, H, L) from a DiffNet state dict (keys relative to `denoise_fn.`)."""
    C, M, _ = sd["input_projection.weight"].shape
    H = sd["residual_layers.0.conditioner_projection.weight"].shape[1]
    L = 0
    while ("residual_layers.%d.dilated_conv.weight" % L) in sd:
        L += 1
    return M, C, H, L


def step_embedding(sd, t, dtype=torch.float32):
    """net.py:124-125: mlp(sinusoid(t)) -> [B, C]."""
    C = sd["input_projection.weight"].shape[0]
    e = sinusoidal_pos_emb(t, C).to(dtype)
    e = F.linear(e, sd["mlp.0.weight"].to(dtype), sd["mlp.0.bias"].to(dtype))
    e = mish(e)
    return F.linear(e, sd["mlp.2.weight"].to(dtype), sd["mlp.2.bias"].to(dtype))


def residual_block(sd, l, x, cond, demb, dilation, dtype=torch.float32):
    """net.py:66-84.  x [B,C,T], cond [B,H,T], demb [B,C] -> (x', skip)."""
    p = "residual_layers.%d." % l
    g = lambda k: sd[p + k].to(dtype)
    d = F.linear(demb, g("diffusion_projection.weight"), g("diffusion_projection.bias")).unsqueeze(-1)
    c = F.conv1d(cond, g("conditioner_projection.weight"), g("conditioner_projection.bias"))
    y = x + d
    # zero padding applies to (x + d): the pad region sees 0, not d (net.py:69-71)
    y = F.conv1d(y, g("dilated_conv.weight"), g("dilated_conv.bias"), padding=dilation, dilation=dilation) + c
    gate, filt = torch.chunk(y, 2, dim=1)          # first half = gate (sigmoid), second = filter (tanh)
    y = torch.sigmoid(gate) * torch.tanh(filt)
    y = F.conv1d(y, g("output_projection.weight"), g("output_projection.bias"))
    residual, skip = torch.chunk(y, 2, dim=1)
    return (x + residual) / math.sqrt(2.0), skip


def diffnet_forward(sd, spec, t, cond, dilation_cycle=4, dtype=torch.float32):
    """net.py:112-135.  spec [B,1,M,T], t int64 [B], cond [B,H,T] -> [B,1,M,T]."""
    M, C, H, L = diffnet_dims(sd)
    g = lambda k: sd[k].to(dtype)
    x = spec[:, 0].to(dtype)
    cond = cond.to(dtype)
    x = F.relu(F.conv1d(x, g("input_projection.weight"), g("input_projection.bias")))
    demb = step_embedding(sd, t, dtype)
    skips = []
    for l in range(L):
        x, s = residual_block(sd, l, x, cond, demb, 2 ** (l % dilation_cycle), dtype)
        skips.append(s)
    x = torch.sum(torch.stack(skips), dim=0) / math.sqrt(L)
    x = F.relu(F.conv1d(x, g("skip_projection.weight"), g("skip_projection.bias")))
    x = F.conv1d(x, g("output_projection.weight"), g("output_projection.bias"))
    return x[:, None, :, :]


# --------------------------------------------------------------------------------------
# Gaussian diffusion schedules and samplers (network/diff/diffusion.py)
# --------------------------------------------------------------------------------------

def linear_beta_schedule(timesteps, max_beta=0.02):
    """diffusion.py:40-45 (max_beta must be passed: the reference freezes its default at import)."""
    return np.linspace(1e-4, max_beta, timesteps)


def cosine_beta_schedule(timesteps, s=0.008):
    """diffusion.py:48-58."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)


def make_schedule(betas):
    """diffusion.py:88-120: the 12 registered fp32 buffers, computed in float64 numpy then cast."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1. - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1., ac[:-1])
    pv = betas * (1. - ac_prev) / (1. - ac)
    vals = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1. - ac),
        "log_one_minus_alphas_cumprod": np.log(1. - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1. / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1. / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1. - ac),
        "posterior_mean_coef2": (1. - ac_prev) * np.sqrt(alphas) / (1. - ac),
    }
    return {k: torch.tensor(v, dtype=torch.float32) for k, v in vals.items()}


def _extract(a, t, ndim):
    """diffusion.py:28-31."""
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


def p_sample(sd, sched, x, t, cond, noise, dilation_cycle=4, dtype=torch.float32, clip_denoised=True):
    """One ancestral DDPM step, diffusion.py:146-163.  `noise` replaces the reference's randn draw."""
    S = lambda k: _extract(sched[k].to(dtype), t, x.dim())
    eps = diffnet_forward(sd, x, t, cond, dilation_cycle, dtype)
    x_recon = S("sqrt_recip_alphas_cumprod") * x - S("sqrt_recipm1_alphas_cumprod") * eps      # :131-135
    if clip_denoised:
        x_recon = x_recon.clamp(-1., 1.)                                                       # :150-151
    mean = S("posterior_mean_coef1") * x_recon + S("posterior_mean_coef2") * x                 # :137-141
    logvar = S("posterior_log_variance_clipped")
    nonzero = (1 - (t == 0).to(dtype)).reshape(x.shape[0], *((1,) * (x.dim() - 1)))           # :162
    return mean + nonzero * (0.5 * logvar).exp() * noise.to(dtype)


def plms_x_pred(sched, x, eps, t, interval, dtype=torch.float32):
    """get_x_pred, diffusion.py:171-179."""
    ac = sched["alphas_cumprod"].to(dtype)
    a_t = _extract(ac, t, x.dim())
    a_prev = _extract(ac, torch.clamp(t - interval, min=0), x.dim())
    a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
    x_delta = (a_prev - a_t) * ((1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x
                                - 1 / (a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * eps)
    return x + x_delta


def p_sample_plms(sd, sched, x, t, interval, cond, noise_list, dilation_cycle=4, dtype=torch.float32):
    """One PLMS step, diffusion.py:166-198.  `noise_list` is the caller-owned eps history (maxlen 4)."""
    eps = diffnet_forward(sd, x, t, cond, dilation_cycle, dtype)
    if len(noise_list) == 0:
        x_pred = plms_x_pred(sched, x, eps, t, interval, dtype)
        eps_prev = diffnet_forward(sd, x_pred, torch.clamp(t - interval, min=0), cond, dilation_cycle, dtype)
        eps_prime = (eps + eps_prev) / 2
    elif len(noise_list) == 1:
        eps_prime = (3 * eps - noise_list[-1]) / 2
    elif len(noise_list) == 2:
        eps_prime = (23 * eps - 16 * noise_list[-1] + 5 * noise_list[-2]) / 12
    else:
        eps_prime = (55 * eps - 59 * noise_list[-1] + 37 * noise_list[-2] - 9 * noise_list[-3]) / 24
    x_prev = plms_x_pred(sched, x, eps_prime, t, interval, dtype)
    noise_list.append(eps)
    if len(noise_list) > 4:
        del noise_list[0]
    return x_prev


def sample(sd, sched, cond, x, t_start, noises=None, pndm_speedup=1, dilation_cycle=4,
           dtype=torch.float32, progress=None):
    """The sampling loop of GaussianDiffusion.forward, diffusion.py:269-278.

    x: initial [B,1,M,T]; noises: [t_start,B,1,M,T] in consumption order (DDPM only: one draw per
    step, including the masked one at t==0).  Returns the final x [B,1,M,T]."""
    x = x.to(dtype)
    b = x.shape[0]
    if pndm_speedup and pndm_speedup > 1:
        hist = []
        for i in reversed(range(0, t_start, pndm_speedup)):
            t = torch.full((b,), i, dtype=torch.long)
            x = p_sample_plms(sd, sched, x, t, pndm_speedup, cond, hist, dilation_cycle, dtype)
            if progress:
                progress(i)
    else:
        for n, i in enumerate(reversed(range(0, t_start))):
            t = torch.full((b,), i, dtype=torch.long)
            x = p_sample(sd, sched, x, t, cond, noises[n], dilation_cycle, dtype)
            if progress:
                progress(i)
    return x


def norm_spec(x, spec_min, spec_max):
    """diffusion.py:286-287."""
    return (x - spec_min) / (spec_max - spec_min) * 2 - 1


def denorm_spec(x, spec_min, spec_max):
    """diffusion.py:289-290."""
    return (x + 1) / 2 * (spec_max - spec_min) + spec_min


def q_sample(sched, x_start, t, noise):
    """diffusion.py:200-205."""
    return (_extract(sched["sqrt_alphas_cumprod"], t, x_start.dim()) * x_start
            + _extract(sched["sqrt_one_minus_alphas_cumprod"], t, x_start.dim()) * noise)


def mel_from_x(x, spec_min, spec_max, mel2ph=None):
    """diffusion.py:279-283: x [B,1,M,T] -> mel_out [B,T,M] (denormalised, masked by mel2ph>0)."""
    y = denorm_spec(x[:, 0].transpose(1, 2), spec_min, spec_max)
    if mel2ph is not None:
        y = y * ((mel2ph > 0).float()[:, :, None])
    return y


# --------------------------------------------------------------------------------------
# Conditioning encoder: FastSpeech2.forward with no_fs2 (modules/fastspeech/fs2.py:94-154,185-238)
# --------------------------------------------------------------------------------------

def f0_to_coarse(f0, f0_bin=256, f0_max=1100.0, f0_min=40.0):
    """utils/pitch_utils.py:17-31 (torch branch)."""
    mel_min = 1127 * np.log(1 + f0_min / 700)
    mel_max = 1127 * np.log(1 + f0_max / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    pos = f0_mel > 0
    f0_mel = torch.where(pos, (f0_mel - mel_min) * (f0_bin - 2) / (mel_max - mel_min) + 1, f0_mel)
    f0_mel = torch.where(f0_mel <= 1, torch.ones_like(f0_mel), f0_mel)
    f0_mel = torch.where(f0_mel > f0_bin - 1, torch.full_like(f0_mel, f0_bin - 1), f0_mel)
    return (f0_mel + 0.5).long()


def cond_encoder(pitch_embed_weight, hubert, mel2ph, f0, f0_bin=256, f0_max=1100.0, f0_min=40.0):
    """fs2.py:94-154 with no_fs2=True, use_pitch_embed=True, pitch_norm='log', use_uv=False,
    no speaker / energy embedding (config_nsf.yaml).  hubert [B,Th,H], mel2ph int64 [B,T],
    f0 [B,T] (log2 domain) -> decoder_inp [B,T,H], f0_denorm [B,T]."""
    enc = F.pad(hubert, [0, 0, 1, 0])                                   # fs2.py:132
    idx = mel2ph[..., None].repeat([1, 1, hubert.shape[-1]])
    dec = torch.gather(enc, 1, idx)                                     # :134-135
    nonpad = (mel2ph > 0).float()[:, :, None]                           # :137
    f0_denorm = 2 ** f0                                                 # pitch_utils.py:66-67
    f0_denorm = torch.where(mel2ph == 0, torch.zeros_like(f0_denorm), f0_denorm)   # :74-75
    pitch = f0_to_coarse(f0_denorm, f0_bin, f0_max, f0_min)             # fs2.py:229
    emb = F.embedding(pitch, pitch_embed_weight)                        # :233 (padding_idx only affects grads)
    return (dec + emb) * nonpad, f0_denorm                              # :145


# --------------------------------------------------------------------------------------
# NSF-HiFiGAN generator (modules/nsf_hifigan/models.py, effective Generator at :325)
# --------------------------------------------------------------------------------------

def fold_weight_norm(sd):
    """remove_weight_norm (models.py:389-396): w = g * v / ||v||, norm over all dims but 0.
    Accepts a checkpoint-style dict with `*.weight_g` / `*.weight_v`; returns plain `*.weight`."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[:-len(".weight_g")]
            wv = sd[base + ".weight_v"]
            norm = wv.reshape(wv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base + ".weight"] = wv * (v / norm)
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def get_padding(kernel_size, dilation=1):
    """modules/nsf_hifigan/utils.py:34-35."""
    return int((kernel_size * dilation - dilation) / 2)


def sine_gen(f0, sampling_rate, harmonic_num, rand_ini, noise, sine_amp=0.1, noise_std=0.003,
             voiced_threshold=0.0):
    """SineGen.forward / _f02sine / _f02uv, models.py:177-276.
    f0 [B,L,1] (Hz, 0 = unvoiced); rand_ini [B,dim] replaces torch.rand (:192; column 0 is zeroed
    here as at :194); noise [B,L,dim] replaces randn_like(sine_waves) (:271)."""
    dim = harmonic_num + 1
    mult = torch.arange(1, dim + 1, dtype=f0.dtype, device=f0.device)
    f0_buf = f0 * mult[None, None, :]                                   # :252-257 (f0*(idx+2))
    rad = (f0_buf / sampling_rate) % 1                                  # :188
    ri = rand_ini.clone()
    ri[:, 0] = 0                                                        # :194
    rad = rad.clone()
    rad[:, 0, :] = rad[:, 0, :] + ri                                    # :195
    tmp_over_one = torch.cumsum(rad, 1) % 1                             # :205
    over_idx = (tmp_over_one[:, 1:, :] - tmp_over_one[:, :-1, :]) < 0   # :206-207
    shift = torch.zeros_like(rad)
    shift[:, 1:, :] = over_idx * -1.0                                   # :208-209
    sines = torch.sin(torch.cumsum(rad + shift, dim=1) * 2 * np.pi)     # :211-212
    sine_waves = sines * sine_amp                                       # :260
    uv = (f0 > voiced_threshold).to(f0.dtype)                           # :171-175
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3                # :270
    sine_waves = sine_waves * uv + noise_amp * noise                    # :271-275
    return sine_waves, uv


def source_module(sd, f0_up, sampling_rate, harmonic_num, rand_ini, noise):
    """SourceModuleHnNSF.forward, models.py:310-323 -> har_source [B,L,1]."""
    sine_wavs, uv = sine_gen(f0_up, sampling_rate, harmonic_num, rand_ini, noise)
    return torch.tanh(F.linear(sine_wavs, sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))


def resblock1(sd, prefix, x, kernel_size, dilations):
    """ResBlock1.forward, models.py:57-64."""
    for j, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[prefix + "convs1.%d.weight" % j], sd[prefix + "convs1.%d.bias" % j],
                      dilation=d, padding=get_padding(kernel_size, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[prefix + "convs2.%d.weight" % j], sd[prefix + "convs2.%d.bias" % j],
                      padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def nsf_generator(sd, h, mel, f0, rand_ini, noise, return_source=False):
    """Generator.forward, models.py:361-387 (the second `Generator`, :325, shadows :97).
    sd: weight-norm-folded state dict; h: dict with upsample_rates, upsample_kernel_sizes,
    upsample_initial_channel, resblock_kernel_sizes, resblock_dilation_sizes, sampling_rate.
    mel [B,M,T] (natural-log mel), f0 [B,T] Hz -> wav [B,1,T*prod(rates)]."""
    rates = list(h["upsample_rates"])
    ksizes = list(h["upsample_kernel_sizes"])
    rks = list(h["resblock_kernel_sizes"])
    rds = list(h["resblock_dilation_sizes"])
    hop = int(np.prod(rates))
    har = None
    if f0 is not None:   # modules/hifigan/hifigan.py:145-149 skips the source without f0; the NSF generator always has it
        f0_up = torch.repeat_interleave(f0[:, None], hop, dim=2).transpose(1, 2)     # nearest upsample, :331,:363
        har = source_module(sd, f0_up, h["sampling_rate"], 8, rand_ini, noise).transpose(1, 2)   # [B,1,L]
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)    # :367
    nk = len(rks)
    for i, (u, k) in enumerate(zip(rates, ksizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd["ups.%d.weight" % i], sd["ups.%d.bias" % i], stride=u, padding=(k - u) // 2)
        if har is not None:
            if i + 1 < len(rates):
                s = int(np.prod(rates[i + 1:]))
                xs_ = F.conv1d(har, sd["noise_convs.%d.weight" % i], sd["noise_convs.%d.bias" % i], stride=s, padding=s // 2)
            else:
                xs_ = F.conv1d(har, sd["noise_convs.%d.weight" % i], sd["noise_convs.%d.bias" % i])
            x = x + xs_
        xs = None
        for j in range(nk):
            r = resblock1(sd, "resblocks.%d." % (i * nk + j), x, rks[j], rds[j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)                                                          # default slope 0.01, :383
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    x = torch.tanh(x)
    return (x, har) if return_source else x


def spec2wav(sd, h, mel_log10, f0, rand_ini, noise):
    """NsfHifiGAN.spec2wav_torch, network/vocoders/nsf_hifigan.py:36-45: mel [B,T,M] log10 -> wav [-1]."""
    c = 2.30259 * mel_log10.transpose(2, 1)
    return nsf_generator(sd, h, c, f0, rand_ini, noise).reshape(-1)


# --------------------------------------------------------------------------------------
# Mel analysis and the mel/f0 glue either side of the hot path
# --------------------------------------------------------------------------------------

def slaney_mel_basis(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) as called at modules/nsf_hifigan/nvSTFT.py:87.

    librosa (0.9.1, requirements.txt:39) is a third-party dependency that is absent from /root/reference and
    from this image; this is its published algorithm (Slaney's Auditory Toolbox mel scale: linear below 1 kHz
    at 200/3 Hz per mel, log above with step ln(6.4)/27; triangles between neighbouring band edges; each filter
    scaled by 2 / (right edge - left edge)), vectorised.  PARITY UNPINNED against librosa itself;
    tests/test_mel_analysis.py pins it against torchaudio's independent implementation."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    to_mel = lambda f: np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) / (np.log(6.4) / 27.0),
                                f * 3.0 / 200.0)   # noqa: E731
    to_hz = lambda m: np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), m * 200.0 / 3.0)  # noqa: E731
    edges = to_hz(np.linspace(to_mel(np.float64(fmin)), to_mel(np.float64(fmax)), n_mels + 2))   # [n_mels+2] Hz
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    rising = (freqs[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    falling = (edges[2:, None] - freqs[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling))
    return (tri * (2.0 / (edges[2:] - edges[:-2]))[:, None]).astype(np.float32)


def mel_analysis(y, n_fft, win_size, hop, mel_basis, clip_val=1e-5, dtype=torch.float32):
    """STFT.get_mel, modules/nsf_hifigan/nvSTFT.py:72-104: y [B, n] -> natural-log mel [B, n_mels, frames].

    `mel_basis` [n_mels, n_fft/2+1] is the matrix of :87-88; dtype=torch.float64 is the tighter arbiter."""
    y = y.to(dtype)
    basis = torch.as_tensor(mel_basis).to(dtype)
    window = torch.hann_window(win_size).to(dtype)                                       # :89
    p = int((n_fft - hop) / 2)
    y = F.pad(y.unsqueeze(1), (p, p), mode="reflect").squeeze(1)                         # :91-92
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win_size, window=window, center=False,
                      normalized=False, onesided=True, return_complex=True)              # :94-95
    spec = torch.view_as_real(spec)
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)                                        # :97
    spec = torch.matmul(basis, spec)                                                     # :99
    return torch.log(torch.clamp(spec, min=clip_val))                                    # :101, :55-56


def wav2spec(wav, n_fft, win_size, hop, mel_basis, dtype=torch.float32):
    """NsfHifiGAN.wav2spec, network/vocoders/nsf_hifigan.py:87-91: wav [n] -> log10 mel [frames, n_mels]."""
    mel = mel_analysis(wav.unsqueeze(0), n_fft, win_size, hop, mel_basis, dtype=dtype).squeeze(0).T
    return (0.434294 * mel) if dtype == torch.float32 else mel * 0.434294


def after_infer_frames(mel_pred, f0_pred, vmin, vmax):
    """The array half of Svc.after_infer, infer_tools/infer_tool.py:182-193 (numpy, B = 1 squeezed):
    mel_pred [T, M], f0_pred [T] -> (clipped kept mel [N, M], kept f0 [N])."""
    mask = np.abs(mel_pred).sum(-1) > 0
    return np.clip(mel_pred[mask], vmin, vmax), f0_pred[mask]


# --------------------------------------------------------------------------------------
# PitchExtractor (modules/fastspeech/pe.py) -- SURVEY.md section 8f row 4
# --------------------------------------------------------------------------------------

def sinusoid_position_table(rows, dim):
    """SinusoidalPositionalEmbedding.get_embedding(rows, dim, padding_idx=0), modules/commons/common_layers.py:105-122."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    ang = torch.arange(rows, dtype=torch.float).unsqueeze(1) * torch.exp(torch.arange(half, dtype=torch.float) * -k).unsqueeze(0)
    tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(rows, -1)
    tab[0, :] = 0
    return tab


def pitch_extractor(sd, mel, conv_layers=2, pad_same=True, pitch_norm="log", use_uv=False, pitch_type="frame",
                    f0_mean=0.0, f0_std=1.0, dtype=torch.float32):
    """PitchExtractor.forward, modules/fastspeech/pe.py:137-149, eval mode.  sd: the module's state_dict;
    mel [B, T, n_mel] -> (pitch_pred [B, T, 2], f0_denorm_pred [B, T])."""
    g = lambda k: sd[k].to(dtype)   # noqa: E731
    mel = mel.to(dtype)
    # Prenet, pe.py:23-44
    nonpad = 1 - mel.abs().sum(-1).eq(0).to(dtype)[:, None, :]                       # [B, 1, T]
    x = mel.transpose(1, 2)
    l = 0
    while "mel_prenet.layers.%d.0.weight" % l in sd:
        p = "mel_prenet.layers.%d." % l
        k = sd[p + "0.weight"].shape[-1]
        x = F.relu(F.conv1d(x, g(p + "0.weight"), g(p + "0.bias"), padding=k // 2))
        x = F.batch_norm(x, g(p + "2.running_mean"), g(p + "2.running_var"), g(p + "2.weight"), g(p + "2.bias"),
                         training=False, eps=1e-5)
        x = x * nonpad
        l += 1
    x = F.linear(x.transpose(1, 2), g("mel_prenet.out_proj.weight"), g("mel_prenet.out_proj.bias"))
    x = x * nonpad.transpose(1, 2)                                                   # [B, T, H]
    # ConvStacks, pe.py:100-117
    if conv_layers > 0:
        x = F.linear(x, g("mel_encoder.in_proj.weight"), g("mel_encoder.in_proj.bias")).transpose(1, -1)
        for i in range(conv_layers):
            p = "mel_encoder.conv.%d." % i
            k = sd[p + "conv.conv.weight"].shape[-1]
            c = F.conv1d(x, g(p + "conv.conv.weight"), g(p + "conv.conv.bias"), padding=k // 2)
            c = F.group_norm(c, sd[p + "norm.weight"].numel() // 16, g(p + "norm.weight"), g(p + "norm.bias"), eps=1e-5)
            x = x + F.relu(c)
        x = F.linear(x.transpose(1, -1), g("mel_encoder.out_proj.weight"), g("mel_encoder.out_proj.bias"))
    # PitchPredictor, tts_modules.py:222-235
    B, T, H = x.shape
    tok = x[..., 0].ne(0).int()                                                      # utils/__init__.py:154-157
    pos = (torch.cumsum(tok, dim=1).type_as(tok) * tok).long()
    table = sinusoid_position_table(max(4096, T + 1), H).to(dtype)
    x = x + g("pitch_predictor.pos_embed_alpha") * table.index_select(0, pos.view(-1)).view(B, T, -1)
    x = x.transpose(1, -1)
    i = 0
    while "pitch_predictor.conv.%d.1.weight" % i in sd:
        p = "pitch_predictor.conv.%d." % i
        k = sd[p + "1.weight"].shape[-1]
        x = F.pad(x, ((k - 1) // 2, (k - 1) // 2) if pad_same else (k - 1, 0))
        x = F.relu(F.conv1d(x, g(p + "1.weight"), g(p + "1.bias")))
        x = F.layer_norm(x.transpose(1, -1), (x.shape[1],), g(p + "3.weight"), g(p + "3.bias"), eps=1e-12).transpose(1, -1)
        i += 1
    pred = F.linear(x.transpose(1, -1), g("pitch_predictor.linear.weight"), g("pitch_predictor.linear.bias"))
    # denorm_f0, utils/pitch_utils.py:63-76
    f0 = pred[:, :, 0].clone()
    if pitch_norm == "standard":
        f0 = f0 * f0_std + f0_mean
    if pitch_norm == "log":
        f0 = 2 ** f0
    if pitch_type == "frame" and use_uv:
        f0[pred[:, :, 1] > 0] = 0
    f0[mel.abs().sum(-1) == 0] = 0
    return pred, f0


# --------------------------------------------------------------------------------------
# Deterministic synthetic parameters live in the neutral module `synthetic.py` (repo root); re-exported
# here because the parity tests spell them `O.synth_*`.
# --------------------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthetic import NSF_H_44K, synth_diffnet_weights, synth_nsf_weights, synth_f0  # noqa: E402,F401
