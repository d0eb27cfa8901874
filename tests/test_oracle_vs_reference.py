"""Live re-check of the oracle against the reference's own modules at the FULL 44.1 kHz config.
Only runs where /root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest
import torch

import ref_harness as rh
from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def hp():
    return rh.install()


def test_full_diffnet_and_ddpm_steps(hp):
    diffusion, net = rh.import_diffusion()
    torch.manual_seed(3)
    dn = net.DiffNet(128).eval()
    torch.nn.init.normal_(dn.output_projection.weight, std=0.05)
    gd = diffusion.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, loss_type="l2",
                                     spec_min=hp["spec_min"], spec_max=hp["spec_max"]).eval()
    sd = {k: v for k, v in dn.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 1, 128, 96, generator=g)
    cond = torch.randn(1, 256, 96, generator=g) * 0.5
    sched = {k: getattr(gd, k) for k in O.SCHEDULE_KEYS}
    for tt in (999, 500, 0):
        t = torch.tensor([tt])
        with torch.no_grad():
            ref = dn(x, t, cond)
        assert (O.diffnet_forward(sd, x, t, cond) - ref).abs().max().item() <= 5e-6
        noise = torch.randn(1, 1, 128, 96, generator=g)
        orig = diffusion.noise_like
        diffusion.noise_like = lambda shape, device, repeat=False: noise
        try:
            with torch.no_grad():
                ref_x = gd.p_sample(x, t, cond)
        finally:
            diffusion.noise_like = orig
        assert (O.p_sample(sd, sched, x, t, cond, noise) - ref_x).abs().max().item() <= 5e-6


def test_synth_weight_keys_match_reference(hp):
    diffusion, net = rh.import_diffusion()
    dn = net.DiffNet(128)
    mine = O.synth_diffnet_weights()
    ref = dn.state_dict()
    assert set(mine) == set(ref)
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k


def test_full_nsf_generator(hp):
    models = rh.import_nsf_models()
    from modules.nsf_hifigan.env import AttrDict
    h = AttrDict(O.NSF_H_44K)
    torch.manual_seed(8)
    gen = models.Generator(h).eval()
    gen.remove_weight_norm()
    mine = O.synth_nsf_weights(O.NSF_H_44K)
    assert set(mine) == set(gen.state_dict())
    gen.load_state_dict(mine)
    T = 6
    g = torch.Generator().manual_seed(2)
    mel = torch.randn(1, 128, T, generator=g) * 2 - 5
    f0 = O.synth_f0(1, T) + 100
    f0[0, 2] = 0
    L = T * 512
    rand_ini = torch.rand(1, 9, generator=g)
    noise = torch.randn(1, L, 9, generator=g)
    draws = iter([rand_ini, noise, torch.zeros(1, L, 1)])
    o_rand, o_randn_like = torch.rand, torch.randn_like
    torch.rand = lambda *a, **k: next(draws).clone()
    torch.randn_like = lambda *a, **k: next(draws).clone()
    try:
        with torch.no_grad():
            ref = gen(mel, f0)
    finally:
        torch.rand, torch.randn_like = o_rand, o_randn_like
    wav = O.nsf_generator(mine, O.NSF_H_44K, mel, f0, rand_ini, noise)
    assert wav.shape == ref.shape == (1, 1, L)
    assert float(ref.std()) > 1e-3                      # not vacuous
    assert (wav - ref).abs().max().item() <= 5e-6


def test_full_pitch_extractor(hp):
    """PitchExtractor at the real size (hidden 256, 80 mel bins), eval mode, perturbed BatchNorm statistics."""
    import modules.fastspeech.pe as pe_mod
    torch.manual_seed(21)
    m = pe_mod.PitchExtractor(n_mel_bins=80, conv_layers=2).eval()
    with torch.no_grad():
        for name, b in m.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.2 * torch.randn_like(b))
            if name.endswith("running_var"):
                b.copy_(0.5 + torch.rand_like(b))
        m.pitch_predictor.linear.bias.add_(torch.tensor([7.0, 0.0]))
    mel = torch.randn(2, 120, 80) - 3.0
    mel[1, 100:] = 0
    with torch.no_grad():
        ret = m(mel)
    pred, f0 = O.pitch_extractor(m.state_dict(), mel)
    assert (pred - ret["pitch_pred"]).abs().max().item() <= 1e-5
    assert (f0 - ret["f0_denorm_pred"]).abs().max().item() <= 1e-3
    assert (f0[1, 100:] == 0).all()


def test_full_mel_analysis(hp):
    """STFT.get_mel at config_nsf.yaml's analysis parameters (2048 / 512 / 128 bins, 40-16000 Hz).  librosa is
    absent: the harness serves librosa.filters.mel from the oracle's restatement (see tests/golden/make_golden.py)."""
    import modules.nsf_hifigan.nvSTFT as nv
    nv.librosa_mel_fn = lambda sr, n_fft, n_mels, fmin, fmax: O.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    stft_now = torch.stft
    nv.torch.stft = lambda *a, **k: stft_now(*a, **k) if "return_complex" in k else torch.view_as_real(stft_now(*a, return_complex=True, **k))
    try:
        g = torch.Generator().manual_seed(4)
        wav = (torch.rand(1, 30000, generator=g) * 2 - 1) * 0.3
        stft = nv.STFT(hp["audio_sample_rate"], hp["audio_num_mel_bins"], hp["fft_size"], hp["win_size"], hp["hop_size"],
                       hp["fmin"], hp["fmax"])
        with torch.no_grad():
            ref = stft.get_mel(wav)
    finally:
        torch.stft = stft_now
    basis = O.slaney_mel_basis(hp["audio_sample_rate"], hp["fft_size"], hp["audio_num_mel_bins"], hp["fmin"], hp["fmax"])
    got = O.mel_analysis(wav, hp["fft_size"], hp["win_size"], hp["hop_size"], basis)
    assert got.shape == ref.shape == (1, 128, 30000 // 512)
    assert (got - ref).abs().max().item() <= 1e-5
