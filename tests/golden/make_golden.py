"""Generate golden input/output vectors from the UNMODIFIED reference (CPU, fp32).

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports the reference's own `DiffNet`, `GaussianDiffusion` (incl. its FastSpeech2
conditioning) and NSF-HiFiGAN `Generator`, builds them at SMALL sizes with seeded
weights, records every random draw the reference makes (torch.randn / rand /
randn_like are wrapped), and writes weights + inputs + draws + outputs to
`tests/golden/*.npz`.  The reference has no tests of its own for this path
(SURVEY.md section 4), so these fixtures are what pins `oracle/diffsvc_oracle.py`.
The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys
from collections import deque

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

SMALL = dict(hidden_size=32, residual_layers=4, residual_channels=64, dilation_cycle_length=2,
             audio_num_mel_bins=16, keep_bins=16)


class DrawRecorder:
    """Wraps torch.randn / rand / randn_like so every draw of the reference is logged in order."""

    def __init__(self):
        self.log = []

    def __enter__(self):
        self._orig = (torch.randn, torch.rand, torch.randn_like)
        o_randn, o_rand, o_randn_like = self._orig

        def randn(*a, **k):
            r = o_randn(*a, **k); self.log.append(("randn", r.clone())); return r

        def rand(*a, **k):
            r = o_rand(*a, **k); self.log.append(("rand", r.clone())); return r

        def randn_like(*a, **k):
            r = o_randn_like(*a, **k); self.log.append(("randn_like", r.clone())); return r

        torch.randn, torch.rand, torch.randn_like = randn, rand, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand, torch.randn_like = self._orig


def _np(sd, prefix="sd/"):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def build_small_diffusion(hp, spec_min, spec_max, K_step, seed):
    diffusion, net = rh.import_diffusion()
    torch.manual_seed(seed)
    dn = net.DiffNet(hp["audio_num_mel_bins"])
    torch.nn.init.normal_(dn.output_projection.weight, std=0.05)   # zero-init in the reference (net.py:110)
    gd = diffusion.GaussianDiffusion(None, hp["audio_num_mel_bins"], dn, timesteps=hp["timesteps"], K_step=K_step,
                                     loss_type=hp["diff_loss_type"], spec_min=spec_min, spec_max=spec_max)
    gd.eval()
    return gd, diffusion


def synth_inputs(B, T, Th, H, seed):
    g = torch.Generator().manual_seed(seed)
    hubert = torch.randn(B, Th, H, generator=g)
    mel2ph = torch.randint(1, Th + 1, (B, T), generator=g).sort(dim=1).values
    mel2ph[-1, T - 3:] = 0                       # padded tail on the last item
    f0 = torch.log2(torch.rand(B, T, generator=g) * 500 + 60)
    f0[0, 2:5] = torch.log2(torch.tensor(1300.0))  # above f0_max -> coarse bin clamp
    return hubert, mel2ph, f0


def gen_diffnet(hp):
    diffusion, net = rh.import_diffusion()
    torch.manual_seed(101)
    dn = net.DiffNet(hp["audio_num_mel_bins"]).eval()
    torch.nn.init.normal_(dn.output_projection.weight, std=0.05)
    g = torch.Generator().manual_seed(5)
    B, M, T, H = 2, hp["audio_num_mel_bins"], 24, hp["hidden_size"]
    spec = torch.randn(B, 1, M, T, generator=g)
    cond = torch.randn(B, H, T, generator=g) * 0.5
    t = torch.tensor([17, 503], dtype=torch.long)
    with torch.no_grad():
        out = dn(spec, t, cond)
    d = _np(dn.state_dict())
    d.update(spec=spec.numpy(), cond=cond.numpy(), t=t.numpy(), out=out.numpy(),
             dilation_cycle=np.int64(hp["dilation_cycle_length"]))
    np.savez_compressed(os.path.join(HERE, "diffnet_small.npz"), **d)
    print("diffnet_small: out", tuple(out.shape), float(out.abs().max()))


def gen_sampler(hp, name, K_step, speedup, spec_min, spec_max, use_gt_mel=False, add_noise_step=0, seed=7):
    gd, diffusion = build_small_diffusion(hp, spec_min, spec_max, K_step, seed)
    B, T, Th, H, M = 2, 20, 9, hp["hidden_size"], hp["audio_num_mel_bins"]
    hubert, mel2ph, f0 = synth_inputs(B, T, Th, H, seed + 1)
    ref_mels = torch.randn(B, T, M, generator=torch.Generator().manual_seed(seed + 2)) * 0.8 - 2.5
    hp["pndm_speedup"] = speedup
    kwargs = dict(use_gt_mel=True, add_noise_step=add_noise_step) if use_gt_mel else {}
    if speedup > 1:
        # the reference's PLMS path only works for B=1 (diffusion.py:186 python max on a tensor);
        # run items one at a time and stack -- per-item semantics
        outs, draws = [], []
        for b in range(B):
            gd.noise_list = deque(maxlen=4)
            with DrawRecorder() as rec, torch.no_grad():
                ret = gd(hubert[b:b + 1], mel2ph[b:b + 1], None, ref_mels[b:b + 1], f0[b:b + 1].clone(), None, None,
                         infer=True, **kwargs)
            outs.append(ret)
            draws.append([r for _, r in rec.log])
        mel_out = torch.cat([o["mel_out"] for o in outs])
        dec = torch.cat([o["decoder_inp"] for o in outs])
        f0d = torch.cat([o["f0_denorm"] for o in outs])
        x_init = torch.cat([d[0] for d in draws])
        noises = np.zeros((0,), np.float32)
    else:
        with DrawRecorder() as rec, torch.no_grad():
            ret = gd(hubert, mel2ph, None, ref_mels, f0.clone(), None, None, infer=True, **kwargs)
        mel_out, dec, f0d = ret["mel_out"], ret["decoder_inp"], ret["f0_denorm"]
        log = [r for _, r in rec.log]
        x_init = log[0]                     # diffusion.py:268 (or the q_sample noise, :205)
        noises = torch.stack(log[1:]).numpy()
    # keep only what the inference path reads (the unused pitch_predictor / mel_out heads stay out)
    d = _np({k: v for k, v in gd.state_dict().items()
             if k.startswith("denoise_fn.") or not k.startswith("fs2.") or k == "fs2.pitch_embed.weight"})
    d.update(hubert=hubert.numpy(), mel2ph=mel2ph.numpy(), f0=f0.numpy(), ref_mels=ref_mels.numpy(),
             x_init=x_init.numpy(), noises=noises, mel_out=mel_out.numpy(), decoder_inp=dec.numpy(),
             f0_denorm=f0d.numpy(), K_step=np.int64(K_step), speedup=np.int64(speedup),
             use_gt_mel=np.int64(int(use_gt_mel)), add_noise_step=np.int64(add_noise_step),
             dilation_cycle=np.int64(hp["dilation_cycle_length"]),
             f0_bin=np.int64(hp["f0_bin"]), f0_max=np.float64(hp["f0_max"]), f0_min=np.float64(hp["f0_min"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, ": mel_out", tuple(mel_out.shape), "draws", len(noises), float(mel_out.abs().max()))


def gen_nsf():
    models = rh.import_nsf_models()
    from modules.nsf_hifigan.env import AttrDict
    h = AttrDict(resblock="1", upsample_rates=[4, 2, 2], upsample_kernel_sizes=[8, 4, 4],
                 upsample_initial_channel=32, resblock_kernel_sizes=[3, 7],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]], num_mels=8, sampling_rate=16000)
    torch.manual_seed(55)
    gen = models.Generator(h).eval()
    # the reference's init (std 0.01) gives a ~0 waveform; perturb v/g so the test is not vacuous
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith("weight_v") or (n.endswith(".weight") and "noise_convs" in n):
                p.mul_(1.0 / (p.std() + 1e-8)).mul_(0.35 / np.sqrt(np.prod(p.shape[1:]) / (4 if n.startswith("ups") else 1)))
            if n.endswith("weight_g"):
                p.copy_(p * (1.0 + 0.3 * torch.randn_like(p)))
    ckpt_sd = {k: v.clone() for k, v in gen.state_dict().items()}     # weight_g / weight_v form
    g = torch.Generator().manual_seed(9)
    B, T = 2, 12
    mel = torch.randn(B, 8, T, generator=g)
    f0 = torch.rand(B, T, generator=g) * 300 + 100
    f0[0, 3:6] = 0
    f0[1, 9:] = 0
    gen.remove_weight_norm()
    with DrawRecorder() as rec, torch.no_grad():
        wav = gen(mel, f0)
    kinds = [k for k, _ in rec.log]
    assert kinds == ["rand", "randn_like", "randn_like"], kinds       # models.py:192, :271, :322
    d = _np(ckpt_sd, "ckpt/")
    d.update(_np(gen.state_dict(), "sd/"))
    d.update(mel=mel.numpy(), f0=f0.numpy(), rand_ini=rec.log[0][1].numpy(), sine_noise=rec.log[1][1].numpy(),
             wav=wav.numpy())
    for k in ("upsample_rates", "upsample_kernel_sizes", "resblock_kernel_sizes", "resblock_dilation_sizes"):
        d["h/" + k] = np.asarray(h[k], dtype=np.int64)
    d["h/upsample_initial_channel"] = np.int64(32)
    d["h/num_mels"] = np.int64(8)
    d["h/sampling_rate"] = np.int64(16000)
    np.savez_compressed(os.path.join(HERE, "nsf_small.npz"), **d)
    print("nsf_small: wav", tuple(wav.shape), float(wav.abs().max()), float(wav.std()))


def gen_hifigan24k():
    """The 24 kHz vocoder's generator (modules/hifigan/hifigan.py:104-169), with and without the f0 source."""
    import modules.hifigan.hifigan as hg
    h = dict(resblock="1", upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=16,
             resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]], use_pitch_embed=True,
             audio_sample_rate=24000)
    torch.manual_seed(77)
    gen = hg.HifiGanGenerator(h).eval()
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith("weight_v") or (n.endswith(".weight") and "noise_convs" in n):
                p.mul_(1.0 / (p.std() + 1e-8)).mul_(0.35 / np.sqrt(np.prod(p.shape[1:]) / (4 if n.startswith("ups") else 1)))
            if n.endswith("weight_g"):
                p.copy_(p * (1.0 + 0.3 * torch.randn_like(p)))
    ckpt_sd = {k: v.clone() for k, v in gen.state_dict().items()}
    g = torch.Generator().manual_seed(19)
    B, T = 2, 10
    mel = torch.randn(B, 80, T, generator=g)
    f0 = torch.rand(B, T, generator=g) * 300 + 100
    f0[1, 2:5] = 0
    gen.remove_weight_norm()
    with DrawRecorder() as rec, torch.no_grad():
        wav_f0 = gen(mel, f0)
    assert [k for k, _ in rec.log] == ["rand", "randn_like", "randn_like"]
    with torch.no_grad():
        wav_plain = gen(mel)
    d = _np(ckpt_sd, "ckpt/")
    d.update(mel=mel.numpy(), f0=f0.numpy(), rand_ini=rec.log[0][1].numpy(), sine_noise=rec.log[1][1].numpy(),
             wav_f0=wav_f0.numpy(), wav_plain=wav_plain.numpy())
    for k in ("upsample_rates", "upsample_kernel_sizes", "resblock_kernel_sizes", "resblock_dilation_sizes"):
        d["h/" + k] = np.asarray(h[k], dtype=np.int64)
    d["h/upsample_initial_channel"] = np.int64(16)
    d["h/audio_sample_rate"] = np.int64(24000)
    np.savez_compressed(os.path.join(HERE, "hifigan24k_small.npz"), **d)
    print("hifigan24k_small: wav", tuple(wav_f0.shape), float(wav_f0.std()), float(wav_plain.std()))


def main():
    hp = rh.install(overrides=SMALL)
    gen_diffnet(hp)
    gen_sampler(hp, "ddpm_small", K_step=6, speedup=1, spec_min=[-5.0], spec_max=[0.0])
    per_bin_min = list(np.linspace(-6.0, -4.0, 16))
    per_bin_max = list(np.linspace(-0.5, 0.5, 16))
    gen_sampler(hp, "ddpm_perbin_small", K_step=5, speedup=1, spec_min=per_bin_min, spec_max=per_bin_max, seed=17)
    gen_sampler(hp, "plms_small", K_step=100, speedup=20, spec_min=[-5.0], spec_max=[0.0], seed=27)
    gen_sampler(hp, "gtmel_small", K_step=1000, speedup=1, spec_min=[-5.0], spec_max=[0.0],
                use_gt_mel=True, add_noise_step=5, seed=37)
    gen_nsf()
    gen_hifigan24k()


if __name__ == "__main__":
    main()
