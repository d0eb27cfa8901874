"""`GaussianDiffusion` -- drop-in for network/diff/diffusion.py:67-296 (inference path).

The T-step sampling loop (diffusion.py:269-278) runs entirely inside libdsvc: one call to
`dsvc_sample_ddpm` / `dsvc_sample_plms` replaces `t` python iterations x ~190 eager kernels.
Constructor signature, `forward` kwargs, returned dict keys, registered buffers and state_dict keys
are the reference's.
"""
import ctypes as C
from collections import deque
from functools import partial

import numpy as np
import torch
from torch import nn

from . import _lib
from .hparams import hparams

try:  # drop-in: keep the reference's own conditioning module
    from modules.fastspeech.fs2 import FastSpeech2  # type: ignore
except Exception:
    from .cond import CondEncoder as FastSpeech2


def exists(x):
    return x is not None


def extract(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def linear_beta_schedule(timesteps, max_beta=None):
    """diffusion.py:40-45.  Unlike the reference, max_beta is read at CALL time (the reference freezes
    the default at import); real checkpoints overwrite the buffers anyway."""
    if max_beta is None:
        max_beta = hparams.get("max_beta", 0.01)
    return np.linspace(1e-4, max_beta, timesteps)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


beta_schedule = {"cosine": cosine_beta_schedule, "linear": linear_beta_schedule}


class GaussianDiffusion(nn.Module):
    def __init__(self, phone_encoder, out_dims, denoise_fn, timesteps=1000, K_step=1000,
                 loss_type=None, betas=None, spec_min=None, spec_max=None):
        super().__init__()
        self.denoise_fn = denoise_fn
        self.fs2 = FastSpeech2(phone_encoder, out_dims)
        self.mel_bins = out_dims
        if exists(betas):
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else betas
        elif "schedule_type" in hparams.keys():
            betas = beta_schedule[hparams["schedule_type"]](timesteps)
        else:
            betas = cosine_beta_schedule(timesteps)
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.K_step = K_step
        self.loss_type = loss_type if loss_type is not None else hparams.get("diff_loss_type", "l1")
        self.noise_list = deque(maxlen=4)   # API parity only: the eps history lives in the native handle
        to_torch = partial(torch.tensor, dtype=torch.float32)
        self.register_buffer("betas", to_torch(betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - alphas_cumprod)))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1. - alphas_cumprod)))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1. / alphas_cumprod)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1. / alphas_cumprod - 1)))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer("posterior_variance", to_torch(posterior_variance))
        self.register_buffer("posterior_log_variance_clipped", to_torch(np.log(np.maximum(posterior_variance, 1e-20))))
        self.register_buffer("posterior_mean_coef1", to_torch(betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod)))
        self.register_buffer("posterior_mean_coef2", to_torch((1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod)))
        self.register_buffer("spec_min", torch.FloatTensor(spec_min)[None, None, :hparams["keep_bins"]])
        self.register_buffer("spec_max", torch.FloatTensor(spec_max)[None, None, :hparams["keep_bins"]])
        if getattr(denoise_fn, "num_timesteps", self.num_timesteps) != self.num_timesteps:
            denoise_fn.num_timesteps = self.num_timesteps

    # ---- native sampler -------------------------------------------------------------------
    def _sync_schedule(self, h):
        bufs = [self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1,
                self.posterior_mean_coef2, self.posterior_log_variance_clipped, self.alphas_cumprod]
        key = (h.value,) + tuple((b.data_ptr(), b._version) for b in bufs)
        if getattr(self.denoise_fn, "_sched_key", None) != key:
            host = [b.detach().to("cpu", torch.float32).contiguous() for b in bufs]
            _lib.check(_lib.load().dsvc_diffnet_set_schedule(h, *[_lib.fptr(t) for t in host]))
            self.denoise_fn._sched_key = key

    def sample(self, x, cond, t, pndm_speedup=None, noise=None, lengths=None, seed=None):
        """The loop of diffusion.py:269-278 on the device.  x [B,1,M,T] (consumed), cond [B,H,T].
        noise: optional [t,B,1,M,T] injected N(0,1) draws (DDPM), else the library's Philox stream."""
        with torch.cuda.device(cond.device):      # the handle's kernels launch on the current device: the tensors' own
            return self._sample(x, cond, t, pndm_speedup, noise, lengths, seed)

    def _sample(self, x, cond, t, pndm_speedup, noise, lengths, seed):
        h = self.denoise_fn.prepare(cond, lengths)
        self._sync_schedule(h)
        x = x.detach().to(torch.float32).contiguous().clone()
        lib = _lib.load()
        if pndm_speedup and pndm_speedup > 1:
            _lib.check(lib.dsvc_sample_plms(h, _lib.dptr(x), int(t), int(pndm_speedup), _lib.current_stream()))
        else:
            nptr = None
            if noise is not None:
                noise = noise.detach().to(torch.float32).contiguous()
                assert noise.shape == (int(t),) + tuple(x.shape), (noise.shape, x.shape)
                nptr = _lib.dptr(noise)
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            _lib.check(lib.dsvc_sample_ddpm(h, _lib.dptr(x), int(t), nptr, C.c_uint64(seed), _lib.current_stream()))
            self._noise_keep = noise
        return x

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def forward(self, hubert, mel2ph=None, spk_embed=None, ref_mels=None, f0=None, uv=None, energy=None,
                infer=False, **kwargs):
        """Inference only (`infer=True`).  Extra, optional kwargs beyond the reference's
        `use_gt_mel` / `add_noise_step`: `x_init`, `noise`, `seed` (injected randomness for parity tests)."""
        x_init, noise, seed = kwargs.pop("x_init", None), kwargs.pop("noise", None), kwargs.pop("seed", None)
        ret = self.fs2(hubert, mel2ph, spk_embed, None, f0, uv, energy, skip_decoder=True, infer=infer, **kwargs)
        cond = ret["decoder_inp"].transpose(1, 2)
        b, device = hubert.shape[0], hubert.device
        if not infer:
            raise NotImplementedError("diffsvc_b200 implements the inference hot path; training is out of scope")
        if kwargs.get("use_gt_mel"):
            t = kwargs["add_noise_step"]
            print('===>using ground truth mel as start, please make sure parameter "key==0" !')
            fs2_mels = self.norm_spec(ref_mels).transpose(1, 2)[:, None, :, :]
            x = self.q_sample(x_start=fs2_mels, t=torch.tensor([t - 1], device=device).long(), noise=x_init)
        else:
            t = self.K_step
            shape = (cond.shape[0], 1, self.mel_bins, cond.shape[2])
            x = torch.randn(shape, device=device) if x_init is None else x_init.to(device)
        lengths = None
        if mel2ph is not None and b > 1:
            # per-item semantics: each item's own length is its conv boundary
            valid = (mel2ph > 0)
            lengths = (valid.float().cumsum(1) * valid.float()).argmax(1) + valid.any(1).long()
            lengths = lengths.tolist()
        speedup = hparams.get("pndm_speedup")
        self.noise_list = deque(maxlen=4)
        x = self.sample(x, cond, t, speedup if (speedup and speedup > 1) else None, noise, lengths, seed)
        x = x[:, 0].transpose(1, 2)
        if mel2ph is not None:
            ret["mel_out"] = self.denorm_spec(x) * ((mel2ph > 0).float()[:, :, None])
        else:
            ret["mel_out"] = self.denorm_spec(x)
        return ret

    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min

    def cwt2f0_norm(self, cwt_spec, mean, std, mel2ph):
        return self.fs2.cwt2f0_norm(cwt_spec, mean, std, mel2ph)

    def out2mel(self, x):
        return x


class OfflineGaussianDiffusion(GaussianDiffusion):
    """Name re-exported because training/task/SVC_task.py:6 imports it; training-only in the reference."""

    def forward(self, *args, **kwargs):
        raise NotImplementedError("OfflineGaussianDiffusion is a training-time class (out of scope)")
