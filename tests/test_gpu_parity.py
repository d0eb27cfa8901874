"""Parity of the CUDA path (through the C-ABI of include/dsvc.h) against
  (a) the golden vectors dumped from the unmodified reference (tests/golden/*.npz), and
  (b) the CPU oracle (oracle/diffsvc_oracle.py) on seeded inputs at the full 44.1 kHz config.

Tolerances (fp32 path; BASELINE.json north_star): <= 1e-3 max-abs on the denoised mel, <= 1e-4 RMS
on the waveform.  The fp32-FFMA mode and the 3-pass tcgen05 mode are held to much tighter bounds on
single evaluations."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _hp(**kw):
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear()
    hparams.update(DEFAULTS_44K)
    hparams.update(kw)
    return hparams


def _load_gold(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return z, sd


def _small_model(sd, K_step, math_mode="fp32"):
    import diffsvc_b200 as D
    C, M, _ = sd["denoise_fn.input_projection.weight"].shape
    H = sd["denoise_fn.residual_layers.0.conditioner_projection.weight"].shape[1]
    L = len([k for k in sd if k.endswith("dilated_conv.weight")])
    _hp(hidden_size=H, residual_layers=L, residual_channels=C, dilation_cycle_length=2, audio_num_mel_bins=M,
        keep_bins=M, pndm_speedup=1)
    dn = D.DiffNet(M, math_mode=math_mode)
    gd = D.GaussianDiffusion(None, M, dn, timesteps=1000, K_step=K_step, loss_type="l2",
                             spec_min=sd["spec_min"].reshape(-1).tolist(), spec_max=sd["spec_max"].reshape(-1).tolist())
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("fs2.") for k in missing), missing
    return gd.to(DEV).eval()


# ------------------------------------------------------------------------------- golden (reference) vectors
def test_diffnet_golden():
    import diffsvc_b200 as D
    z, sd = _load_gold("diffnet_small")
    C, M, _ = sd["input_projection.weight"].shape
    _hp(hidden_size=sd["residual_layers.0.conditioner_projection.weight"].shape[1], residual_layers=4,
        residual_channels=C, dilation_cycle_length=int(z["dilation_cycle"]), audio_num_mel_bins=M, keep_bins=M)
    dn = D.DiffNet(M, math_mode="fp32")
    dn.load_state_dict(sd, strict=True)
    dn = dn.to(DEV)
    spec, cond = torch.from_numpy(z["spec"]).to(DEV), torch.from_numpy(z["cond"]).to(DEV)
    ref = torch.from_numpy(z["out"])
    # the reference evaluates both items at different steps; the native API shares t across the batch
    for b, t in enumerate(z["t"].tolist()):
        out = dn(spec[b:b + 1], torch.tensor([t], device=DEV), cond[b:b + 1]).cpu()
        assert (out - ref[b:b + 1]).abs().max().item() <= 1e-5


@pytest.mark.parametrize("name", ["ddpm_small", "ddpm_perbin_small", "gtmel_small"])
def test_ddpm_golden(name):
    z, sd = _load_gold(name)
    gd = _small_model(sd, int(z["K_step"]))
    hubert, mel2ph, f0 = (torch.from_numpy(z[k]).to(DEV) for k in ("hubert", "mel2ph", "f0"))
    ret = gd.fs2(hubert, mel2ph, None, None, f0.clone(), None, None, skip_decoder=True, infer=True)
    assert (ret["decoder_inp"].cpu() - torch.from_numpy(z["decoder_inp"])).abs().max().item() <= 1e-5
    cond = ret["decoder_inp"].transpose(1, 2)
    x_init = torch.from_numpy(z["x_init"]).to(DEV)
    if int(z["use_gt_mel"]):
        t0 = int(z["add_noise_step"])
        xs = gd.norm_spec(torch.from_numpy(z["ref_mels"]).to(DEV)).transpose(1, 2)[:, None]
        x = gd.q_sample(xs, torch.tensor([t0 - 1], device=DEV), x_init)
    else:
        t0, x = int(z["K_step"]), x_init
    # lengths=None: the reference's padded-batch semantics (what this golden batch was computed with)
    xf = gd.sample(x, cond, t0, None, torch.from_numpy(z["noises"]).to(DEV), None)
    mel = gd.denorm_spec(xf[:, 0].transpose(1, 2)) * ((mel2ph > 0).float()[:, :, None])
    assert (mel.cpu() - torch.from_numpy(z["mel_out"])).abs().max().item() <= 1e-4


def test_plms_golden_via_forward():
    z, sd = _load_gold("plms_small")
    gd = _small_model(sd, int(z["K_step"]))
    from diffsvc_b200.hparams import hparams
    hparams["pndm_speedup"] = int(z["speedup"])
    ref = torch.from_numpy(z["mel_out"])
    for b in range(ref.shape[0]):
        sl = slice(b, b + 1)
        ret = gd(torch.from_numpy(z["hubert"][sl]).to(DEV), torch.from_numpy(z["mel2ph"][sl]).to(DEV), None, None,
                 torch.from_numpy(z["f0"][sl]).to(DEV), None, None, infer=True,
                 x_init=torch.from_numpy(z["x_init"][sl]))
        assert set(["mel_out", "decoder_inp", "f0_denorm", "mel2ph"]) <= set(ret.keys())
        assert (ret["mel_out"].cpu() - ref[sl]).abs().max().item() <= 1e-4


def test_nsf_golden():
    from diffsvc_b200.vocoders.nsf_models import Generator
    z = np.load(os.path.join(GOLD, "nsf_small.npz"))
    ckpt = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ckpt/")}
    h = {k[2:]: z[k].tolist() for k in z.files if k.startswith("h/")}
    h["resblock"] = "1"
    gen = Generator(h, ckpt, device=DEV)       # weight_g / weight_v form: folds the weight norm itself
    wav = gen(torch.from_numpy(z["mel"]).to(DEV), torch.from_numpy(z["f0"]).to(DEV),
              rand_ini=torch.from_numpy(z["rand_ini"]), sine_noise=torch.from_numpy(z["sine_noise"]))
    ref = torch.from_numpy(z["wav"])
    assert wav.shape == ref.shape
    d = (wav.cpu() - ref)
    assert d.abs().max().item() <= 2e-5 and d.pow(2).mean().sqrt().item() <= 5e-6


# ------------------------------------------------------------------------------- full config vs the oracle
def _full_model(math_mode, K_step=1000, seed=1234):
    import diffsvc_b200 as D
    _hp(pndm_speedup=1)
    sd = O.synth_diffnet_weights(seed=seed)
    dn = D.DiffNet(128, math_mode=math_mode)
    dn.load_state_dict(sd, strict=True)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=K_step, loss_type="l2", spec_min=[-5.0], spec_max=[0.0])
    return gd.to(DEV).eval(), sd


def _inputs(B, T, steps, seed=7):
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, 256, T, generator=g) * 0.5
    x0 = torch.randn(B, 1, 128, T, generator=g)
    noise = torch.randn(steps, B, 1, 128, T, generator=g)
    return cond, x0, noise


@pytest.mark.parametrize("math_mode,tol", [("fp32", 2e-5), ("tc3f16", 1e-4)])
def test_diffnet_eval_full(math_mode, tol):
    gd, sd = _full_model(math_mode)
    cond, x0, _ = _inputs(2, 200, 1)
    for t in (999, 37, 0):
        ref = O.diffnet_forward(sd, x0, torch.tensor([t, t]), cond)
        out = gd.denoise_fn(x0.to(DEV), torch.tensor([t, t], device=DEV), cond.to(DEV)).cpu()
        err = (out - ref).abs().max().item()
        assert err <= tol, (math_mode, t, err)


@pytest.mark.parametrize("math_mode", ["fp32", "tc3f16"])
def test_ddpm_chain_full(math_mode):
    steps, T = 60, 136
    gd, sd = _full_model(math_mode)
    cond, x0, noise = _inputs(1, T, steps)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    ref = O.mel_from_x(O.sample(sd, sched, cond, x0, steps, noise), torch.tensor([[[-5.0]]]), torch.tensor([[[0.0]]]))
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV))
    mel = gd.denorm_spec(xf[:, 0].transpose(1, 2)).cpu()
    err = (mel - ref).abs().max().item()
    assert err <= 1e-3, (math_mode, err)          # the north_star gate
    assert err <= (1e-4 if math_mode == "fp32" else 3e-4), (math_mode, err)


@pytest.mark.parametrize("math_mode", ["fp32", "tc3f16"])
def test_plms_chain_full(math_mode):
    gd, sd = _full_model(math_mode)
    cond, x0, _ = _inputs(1, 100, 1, seed=9)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    ref = O.sample(sd, sched, cond, x0, 1000, None, pndm_speedup=100)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), 1000, 100).cpu()
    # with random weights the 10-iteration PLMS solve is expansive (|x| reaches several hundred in
    # normalised units; no clamp in PLMS, diffusion.py:166-198), so the bound is relative to the range:
    # fp32-class agreement = a few 1e-6 of max|x|
    rng = max(1.0, ref.abs().max().item())
    err = (xf - ref).abs().max().item() / rng
    assert err <= 1e-5, (math_mode, err, rng)


@pytest.mark.parametrize("math_mode,pack", [("fp32", "1"), ("tc3f16", "1"), ("tc3f16", "0")])
def test_ragged_batch_is_per_item(math_mode, pack, monkeypatch):
    """Each item of a ragged batch is computed as if alone: the oracle is a loop of B=1 calls.
    tc3f16 batches run PACKED by default (items back to back on one frame axis, max-dilation zero rows between them);
    DSVC_PACK=0 keeps the per-item [B][Tmax] layout with its live-tile table."""
    monkeypatch.setenv("DSVC_PACK", pack)
    steps, lens = 12, [150, 97, 33]
    T = max(lens)
    gd, sd = _full_model(math_mode)
    cond, x0, noise = _inputs(len(lens), T, steps, seed=21)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    for b, n in enumerate(lens):
        ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
        err = (xf[b:b + 1, :, :, :n] - ref).abs().max().item()
        assert err <= 2e-4, (math_mode, b, err)


def test_batch_composition_invariance(monkeypatch):
    """Size-independent property at the BASELINE shape [B,128,1000]: an item's result does not depend on its batch.

    Within one tile class (same tile width, CTA pairs or not, no split-K) every contraction keeps one K-chain and one
    summation order whatever the grid is -> bit-identical.  The launcher picks the tile width from the grid size
    (waves x cost, tc_gemm.cuh), and the classes differ in the order in which the three partial products of the
    error-compensated fp16 split are added -> across classes equal to fp32 rounding, and still deterministic."""
    cond, x0, noise = _inputs(3, 1000, 3, seed=5)

    def run(gd):
        full = gd.sample(x0.to(DEV), cond.to(DEV), 3, None, noise.to(DEV)).cpu()
        one = gd.sample(x0[1:2].to(DEV), cond[1:2].to(DEV), 3, None, noise[:, 1:2].contiguous().to(DEV)).cpu()
        again = gd.sample(x0.to(DEV), cond.to(DEV), 3, None, noise.to(DEV)).cpu()
        assert torch.equal(full, again)            # deterministic
        return full, one

    monkeypatch.delenv("DSVC_SPLITK", raising=False)
    for bn in ("64", "128"):                       # one tile class for both grids
        monkeypatch.setenv("DSVC_TC_BN", bn)
        gd0, sd = _full_model("tc3f16")
        full0, one0 = run(gd0)
        assert torch.equal(full0[1:2], one0), bn
    monkeypatch.delenv("DSVC_TC_BN")
    gd1, _ = _full_model("tc3f16")                 # the launcher's own choice: 128-wide tiles for 3 clips, 64-wide for 1
    full1, one1 = run(gd1)
    assert (full1[1:2] - one1).abs().max().item() <= 2e-5 * one1.abs().max().item()
    monkeypatch.setenv("DSVC_SPLITK", "1")         # opt-in split-K conv of small grids: another order again
    gd2, _ = _full_model("tc3f16")
    full2, one2 = run(gd2)
    assert torch.equal(full2, full1)               # the 3-clip grid is too large for split-K: same kernels
    assert not torch.equal(one2, one1)             # the single clip did take the split-K kernel ...
    assert (one2 - one1).abs().max().item() <= 2e-5 * one1.abs().max().item()     # ... same math, other order
    one2b = gd2.sample(x0[1:2].to(DEV), cond[1:2].to(DEV), 3, None, noise[:, 1:2].contiguous().to(DEV)).cpu()
    assert torch.equal(one2, one2b)


def test_philox_stream_is_seeded():
    """Library-generated noise: deterministic per seed, different across seeds, masked at t == 0."""
    gd, sd = _full_model("fp32")
    cond, x0, _ = _inputs(1, 256, 1)
    run = lambda t, seed: gd.sample(x0.to(DEV), cond.to(DEV), t, None, None, seed=seed).cpu()
    assert torch.equal(run(1, 11), run(1, 12))          # only t=0 runs: the draw is multiplied by 0
    a, b, c = run(2, 11), run(2, 11), run(2, 12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # sigma_1 * (n_a - n_c) propagated through one more (t=0) step: finite, non-trivial spread
    d = (a - c).flatten()
    assert torch.isfinite(d).all() and d.std().item() > 1e-4


def test_nsf_full_vs_oracle():
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    _hp()
    sd = O.synth_nsf_weights(O.NSF_H_44K)
    voc = NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device=DEV)
    B, T = 2, 40
    g = torch.Generator().manual_seed(3)
    mel = torch.randn(B, T, 128, generator=g) * 0.8 - 2.0            # log10 mel
    f0 = O.synth_f0(B, T)
    f0[0, 5:9] = 0
    L = T * 512
    rand_ini = torch.rand(B, 9, generator=g)
    noise = torch.randn(B, L, 9, generator=g)
    ref = O.spec2wav(sd, O.NSF_H_44K, mel, f0, rand_ini, noise)
    wav = voc.spec2wav_torch(mel.to(DEV), f0=f0.to(DEV), rand_ini=rand_ini, sine_noise=noise).cpu()
    assert wav.shape == ref.shape == (B * L,)
    assert float(ref.std()) > 1e-2
    d = wav - ref
    rms, mx = d.pow(2).mean().sqrt().item(), d.abs().max().item()
    assert rms <= 1e-4, (rms, mx)                 # the north_star gate
    assert mx <= 5e-4, (rms, mx)
    # numpy entry point (network/vocoders/nsf_hifigan.py:47-73), item 0
    w0 = voc.spec2wav(mel[0].numpy(), f0=f0[0].numpy(), rand_ini=rand_ini[0:1], sine_noise=noise[0:1])
    assert isinstance(w0, np.ndarray) and w0.dtype == np.float32 and w0.shape == (L,)
    assert np.abs(w0 - ref[:L].numpy()).max() <= 5e-4


@pytest.mark.parametrize("B,T", [(1, 7), (3, 33)])
def test_nsf_tensor_core_path_vs_ffma_path(B, T, monkeypatch):
    """ResBlock convs of the >= 64-channel stages run on tcgen05 (3-pass fp16 split); DSVC_NSF_MATH=fp32 keeps the
    whole generator on the FFMA path.  Both must sit inside the waveform gate, and agree closely."""
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    _hp()
    sd = O.synth_nsf_weights(O.NSF_H_44K, seed=77)
    g = torch.Generator().manual_seed(B * 100 + T)
    mel = torch.randn(B, T, 128, generator=g) * 0.8 - 2.0
    f0 = O.synth_f0(B, T)
    L = T * 512
    rand_ini = torch.rand(B, 9, generator=g)
    noise = torch.randn(B, L, 9, generator=g)
    ref = O.spec2wav(sd, O.NSF_H_44K, mel, f0, rand_ini, noise)
    out = {}
    for mode in ("tc", "fp32"):
        if mode == "fp32":
            monkeypatch.setenv("DSVC_NSF_MATH", "fp32")
        else:
            monkeypatch.delenv("DSVC_NSF_MATH", raising=False)
        voc = NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device=DEV)
        out[mode] = voc.spec2wav_torch(mel.to(DEV), f0=f0.to(DEV), rand_ini=rand_ini, sine_noise=noise).cpu()
        d = out[mode] - ref
        rms = d.pow(2).mean().sqrt().item()
        print("nsf %s B=%d T=%d: rms %.2e max %.2e" % (mode, B, T, rms, d.abs().max().item()))
        assert rms <= 1e-4 and d.abs().max().item() <= 5e-4
    assert (out["tc"] - out["fp32"]).abs().max().item() <= 5e-5


def test_errors_are_loud():
    import diffsvc_b200 as D
    from diffsvc_b200 import _lib
    _hp()
    dn = D.DiffNet(128, math_mode="fp32").to(DEV)
    with pytest.raises(_lib.DsvcError):
        _lib.check(_lib.load().dsvc_diffnet_eval(dn.handle(), None, 0, None, None))   # null args / not prepared
    with pytest.raises(ValueError):
        dn(torch.zeros(2, 1, 128, 8, device=DEV), torch.tensor([1, 2], device=DEV), torch.zeros(2, 256, 8, device=DEV))


# ------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("math_mode", ["fp32", "tc3f16"])
@pytest.mark.parametrize("T", [1, 7, 129])
def test_tiny_and_odd_lengths(math_mode, T):
    """Frame counts below one tile, not a multiple of anything, and one frame past a tile boundary."""
    gd, sd = _full_model(math_mode)
    cond, x0, noise = _inputs(1, T, 3, seed=31)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    ref = O.sample(sd, sched, cond, x0, 3, noise)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), 3, None, noise.to(DEV)).cpu()
    assert (xf - ref).abs().max().item() <= 5e-5


def test_zero_steps_and_empty_item():
    gd, sd = _full_model("tc3f16")
    cond, x0, noise = _inputs(2, 40, 2, seed=3)
    # t_start = 0: the loop body never runs; x comes back unchanged (diffusion.py:276 with t=0)
    same = gd.sample(x0.to(DEV), cond.to(DEV), 0, None, None).cpu()
    assert torch.equal(same, x0)
    # an item of length 0 next to a full one: the full item must equal its stand-alone result
    both = gd.sample(x0.to(DEV), cond.to(DEV), 2, None, noise.to(DEV), lengths=[40, 0]).cpu()
    one = gd.sample(x0[:1].to(DEV), cond[:1].to(DEV), 2, None, noise[:, :1].contiguous().to(DEV)).cpu()
    assert torch.equal(both[:1], one)


def test_plms_interval_not_dividing_t():
    """reversed(range(0, t, interval)) when interval does not divide t (diffusion.py:272)."""
    gd, sd = _full_model("fp32")
    cond, x0, _ = _inputs(1, 48, 1, seed=13)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    ref = O.sample(sd, sched, cond, x0, 250, None, pndm_speedup=60)     # t = 240, 180, 120, 60, 0
    xf = gd.sample(x0.to(DEV), cond.to(DEV), 250, 60).cpu()
    rng = max(1.0, ref.abs().max().item())
    assert (xf - ref).abs().max().item() / rng <= 1e-5


def test_weights_reload_rebuilds_handle():
    """load_state_dict after first use must not keep stale device weights (utils.load_ckpt after .cuda())."""
    gd, sd = _full_model("fp32")
    cond, x0, _ = _inputs(1, 32, 1)
    t = torch.tensor([5], device=DEV)
    a = gd.denoise_fn(x0.to(DEV), t, cond.to(DEV)).cpu()
    sd2 = O.synth_diffnet_weights(seed=99)
    gd.denoise_fn.load_state_dict(sd2)
    b = gd.denoise_fn(x0.to(DEV), t, cond.to(DEV)).cpu()
    ref = O.diffnet_forward(sd2, x0, torch.tensor([5]), cond)
    assert (b - ref).abs().max().item() <= 2e-5 and (a - b).abs().max().item() > 1e-3


def test_vocoder_checkpoint_file_roundtrip(tmp_path):
    """NsfHifiGAN() the reference's way: hparams['vocoder_ckpt'] + sibling config.json + ['generator'] in
    weight_g/weight_v form (modules/nsf_hifigan/models.py:14-30)."""
    import json
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    z = np.load(os.path.join(GOLD, "nsf_small.npz"))
    ckpt = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ckpt/")}
    h = {k[2:]: z[k].tolist() for k in z.files if k.startswith("h/")}
    h.update(resblock="1", n_fft=512, win_size=512, hop_size=16, fmin=40, fmax=8000)
    (tmp_path / "config.json").write_text(json.dumps(h))
    torch.save({"generator": ckpt}, tmp_path / "model")
    _hp(vocoder_ckpt=str(tmp_path / "model"), audio_sample_rate=16000, audio_num_mel_bins=8, hop_size=16, fft_size=512,
        win_size=512, fmin=40, fmax=8000)
    voc = NsfHifiGAN()
    assert voc.h.num_mels == 8 and voc.model.hop == 16
    mel = torch.from_numpy(z["mel"]).transpose(1, 2) / 2.30259           # back to "log10" mel [B,T,M]
    wav = voc.spec2wav_torch(mel.to(DEV), f0=torch.from_numpy(z["f0"]).to(DEV),
                             rand_ini=torch.from_numpy(z["rand_ini"]), sine_noise=torch.from_numpy(z["sine_noise"])).cpu()
    assert (wav - torch.from_numpy(z["wav"]).reshape(-1)).abs().max().item() <= 5e-5


def test_hifigan24k_golden():
    """24 kHz vocoder (network/vocoders/hifigan.py + modules/hifigan/hifigan.py) on the same kernels."""
    from diffsvc_b200.vocoders.hifigan import HifiGAN
    z = np.load(os.path.join(GOLD, "hifigan24k_small.npz"))
    ckpt = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ckpt/")}
    h = {k[2:]: z[k].tolist() for k in z.files if k.startswith("h/")}
    h.update(resblock="1", use_pitch_embed=True)
    _hp(use_nsf=True)
    voc = HifiGAN.from_state_dict(h, ckpt, device=DEV)
    assert voc.model.h.num_mels == 80 and voc.model.hop == 8
    mel, f0 = z["mel"], z["f0"]
    for b in range(mel.shape[0]):
        w = voc.spec2wav(mel[b].T, f0=f0[b], rand_ini=torch.from_numpy(z["rand_ini"][b:b + 1]),
                         sine_noise=torch.from_numpy(z["sine_noise"][b:b + 1]))
        assert np.abs(w - z["wav_f0"][b, 0]).max() <= 2e-5
        w2 = voc.spec2wav(mel[b].T)                                        # no f0: plain HiFi-GAN path
        assert np.abs(w2 - z["wav_plain"][b, 0]).max() <= 2e-5


def test_cond_encoder_kernel_vs_oracle():
    """the fused conditioning kernel against the oracle's fs2 (no_fs2) restatement."""
    from diffsvc_b200.cond import CondEncoder
    _hp()
    enc = CondEncoder().to(DEV)
    g = torch.Generator().manual_seed(4)
    hub = torch.randn(3, 40, 256, generator=g)
    mel2ph = torch.randint(1, 41, (3, 70), generator=g).sort(dim=1).values
    mel2ph[2, 60:] = 0
    f0 = torch.log2(torch.rand(3, 70, generator=g) * 900 + 45)
    f0[0, :4] = torch.log2(torch.tensor(1500.0))                 # above f0_max: clamps to the last bin
    dec, f0d = O.cond_encoder(enc.pitch_embed.weight.detach().cpu(), hub, mel2ph, f0.clone())
    f0_dev = f0.clone().to(DEV)
    ret = enc(hub.to(DEV), mel2ph.to(DEV), None, None, f0_dev, None, None)
    assert (ret["f0_denorm"].cpu() - f0d).abs().max().item() <= 1e-3 * f0d.abs().max().item()
    # a coarse-pitch bin may flip at an exact boundary through 1-ulp log/exp differences: allow none here (seeded)
    assert (ret["decoder_inp"].cpu() - dec).abs().max().item() <= 1e-6
    assert torch.equal(f0_dev.cpu()[2, 60:], torch.zeros(10))   # in-place zeroing of padded f0 (fs2.py:226-227)
