"""The CPU oracle against the golden vectors dumped from the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins oracle/diffsvc_oracle.py: the reference has
no tests or known-answer vectors of its own for this path."""
import os

import numpy as np
import pytest
import torch

from oracle import diffsvc_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    rest = {k: z[k] for k in z.files if "/" not in k}
    return z, sd, rest


def test_diffnet_matches_reference():
    z, sd, r = load("diffnet_small")
    out = O.diffnet_forward(sd, torch.from_numpy(r["spec"]), torch.from_numpy(r["t"]),
                            torch.from_numpy(r["cond"]), int(r["dilation_cycle"]))
    ref = torch.from_numpy(r["out"])
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-6      # same library calls, same order: fp32 round-off only


def _run_sampler(name):
    z, sd, r = load(name)
    dn = {k[len("denoise_fn."):]: v for k, v in sd.items() if k.startswith("denoise_fn.")}
    sched = {k: sd[k] for k in O.SCHEDULE_KEYS}
    hubert, mel2ph, f0 = (torch.from_numpy(r[k]) for k in ("hubert", "mel2ph", "f0"))
    dec, f0d = O.cond_encoder(sd["fs2.pitch_embed.weight"], hubert, mel2ph, f0.clone(),
                              int(r["f0_bin"]), float(r["f0_max"]), float(r["f0_min"]))
    assert (dec - torch.from_numpy(r["decoder_inp"])).abs().max().item() <= 1e-6
    assert (f0d - torch.from_numpy(r["f0_denorm"])).abs().max().item() <= 1e-3   # Hz, 2**f0
    cond = dec.transpose(1, 2)
    K, sp, cyc = int(r["K_step"]), int(r["speedup"]), int(r["dilation_cycle"])
    x_init = torch.from_numpy(r["x_init"])
    if int(r["use_gt_mel"]):
        t0 = int(r["add_noise_step"])
        xs = O.norm_spec(torch.from_numpy(r["ref_mels"]), sd["spec_min"], sd["spec_max"]).transpose(1, 2)[:, None]
        x = O.q_sample(sched, xs, torch.tensor([t0 - 1]), x_init)     # diffusion.py:258-261
    else:
        t0, x = K, x_init
    if sp > 1:
        outs = [O.sample(dn, sched, cond[b:b + 1], x[b:b + 1], t0, None, sp, cyc) for b in range(x.shape[0])]
        xf = torch.cat(outs)
    else:
        xf = O.sample(dn, sched, cond, x, t0, torch.from_numpy(r["noises"]), 1, cyc)
    mel = O.mel_from_x(xf, sd["spec_min"], sd["spec_max"], mel2ph)
    return mel, torch.from_numpy(r["mel_out"])


@pytest.mark.parametrize("name", ["ddpm_small", "ddpm_perbin_small", "plms_small", "gtmel_small"])
def test_sampler_matches_reference(name):
    mel, ref = _run_sampler(name)
    assert mel.shape == ref.shape
    assert (mel - ref).abs().max().item() <= 2e-5


def test_schedule_buffers_match_reference():
    z, sd, r = load("ddpm_small")
    mine = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(mine[k], sd[k]), k


def test_nsf_generator_matches_reference():
    z = np.load(os.path.join(GOLD, "nsf_small.npz"))
    ckpt = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ckpt/")}
    folded_ref = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    h = {k[2:]: (z[k].tolist()) for k in z.files if k.startswith("h/")}
    folded = O.fold_weight_norm(ckpt)
    assert set(folded) == set(folded_ref)
    for k in folded:
        assert (folded[k] - folded_ref[k]).abs().max().item() <= 1e-6, k
    wav = O.nsf_generator(folded, h, torch.from_numpy(z["mel"]), torch.from_numpy(z["f0"]),
                          torch.from_numpy(z["rand_ini"]), torch.from_numpy(z["sine_noise"]))
    ref = torch.from_numpy(z["wav"])
    assert wav.shape == ref.shape
    assert (wav - ref).abs().max().item() <= 2e-6


def test_hifigan24k_generator_matches_reference():
    """modules/hifigan/hifigan.py: same network; with the f0 source and without it (`f0=None`)."""
    z = np.load(os.path.join(GOLD, "hifigan24k_small.npz"))
    ckpt = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ckpt/")}
    h = {k[2:]: (z[k].tolist()) for k in z.files if k.startswith("h/")}
    h["sampling_rate"] = h["audio_sample_rate"]
    sd = O.fold_weight_norm(ckpt)
    mel, f0 = torch.from_numpy(z["mel"]), torch.from_numpy(z["f0"])
    wav = O.nsf_generator(sd, h, mel, f0, torch.from_numpy(z["rand_ini"]), torch.from_numpy(z["sine_noise"]))
    assert (wav - torch.from_numpy(z["wav_f0"])).abs().max().item() <= 2e-6
    plain = O.nsf_generator(sd, h, mel, None, None, None)
    assert (plain - torch.from_numpy(z["wav_plain"])).abs().max().item() <= 2e-6
