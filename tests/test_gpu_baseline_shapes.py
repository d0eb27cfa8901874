"""Parity at the shapes `bench.py` and BASELINE.json's configs actually run, default math (tc3f16 = 3-pass fp16
hi/lo on tcgen05), through the C-ABI, against the CPU oracle on the same seeded inputs and injected noise.

  configs[1]  1000-step DDPM x 862 frames (one 10 s clip): mel <= 1e-3 max-abs AND the waveform vocoded from that
              mel <= 1e-4 RMS against the oracle's own chain (network/diff/diffusion.py:269-278, net.py:66-84,
              modules/nsf_hifigan/models.py:361-387)
  configs[2]  PLMS, interval 40, 862 frames (diffusion.py:166-198)
  configs[4]  PLMS, interval 20, 43 frames (the flask chunk; short-clip kernel selection)
  configs[3]  ragged 8 x ~689 frames, oracle = loop of B=1 calls
  SURVEY 7    the minimum slice: one ResidualBlock at dilation 1, 2, 4, 8 on [1,384,1000] and [4,384,862]
  cosine      the cosine beta schedule (diffusion.py:48-58) through the product constructor

The oracle of configs[1] is ~1-2 min of host CPU; everything else is seconds.
"""
import os
import time

import pytest
import torch

from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
SPEC_MIN, SPEC_MAX = torch.tensor([[[-5.0]]]), torch.tensor([[[0.0]]])


def _hp(**kw):
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear()
    hparams.update(DEFAULTS_44K)
    hparams.update(kw)
    return hparams


def _model(math_mode="tc3f16", K_step=1000, sd=None, **hp):
    import diffsvc_b200 as D
    _hp(pndm_speedup=1, **hp)
    sd = O.synth_diffnet_weights() if sd is None else sd
    dn = D.DiffNet(128, math_mode=math_mode)
    dn.load_state_dict(sd, strict=True)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=K_step, loss_type="l2", spec_min=[-5.0], spec_max=[0.0])
    return gd.to(DEV).eval(), sd


def _inputs(B, T, steps, seed=7):
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, 256, T, generator=g) * 0.5
    x0 = torch.randn(B, 1, 128, T, generator=g)
    noise = torch.randn(steps, B, 1, 128, T, generator=g) if steps else None
    return cond, x0, noise


def _threads():
    torch.set_num_threads(min(32, os.cpu_count() or 1))     # these small convs oversubscribe beyond ~32 threads


def test_cfg1_ddpm1000_x_862_mel_and_waveform():
    """BASELINE configs[1] exactly as benchmarked: 1000 DDPM steps, one 862-frame clip, tc3f16, then NSF-HiFiGAN."""
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    _threads()
    steps, T = 1000, 862
    gd, sd = _model()
    cond, x0, noise = _inputs(1, T, steps)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    t0 = time.time()
    with torch.no_grad():
        ref_mel = O.mel_from_x(O.sample(sd, sched, cond, x0, steps, noise), SPEC_MIN, SPEC_MAX)
    t_oracle = time.time() - t0
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV))
    mel = gd.denorm_spec(xf[:, 0].transpose(1, 2))
    err = (mel.cpu() - ref_mel).abs().max().item()
    # vocoder on each side's own mel (after_infer's clip, infer_tool.py:183): the full-chain waveform error
    nsd = O.synth_nsf_weights(O.NSF_H_44K)
    voc = NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), nsd, device=DEV)
    g = torch.Generator().manual_seed(5)
    f0 = O.synth_f0(1, T)
    rand_ini = torch.rand(1, 9, generator=g)
    sn = torch.randn(1, T * 512, 9, generator=g)
    with torch.no_grad():
        ref_wav = O.spec2wav(nsd, O.NSF_H_44K, ref_mel.clamp(-6.0, 1.5), f0, rand_ini, sn)
    wav = voc.spec2wav_torch(mel.clamp(-6.0, 1.5), f0=f0.to(DEV), rand_ini=rand_ini, sine_noise=sn).cpu()
    rms = (wav - ref_wav).pow(2).mean().sqrt().item()
    print("cfg1 1000x862 tc3f16: mel max-abs %.3e, wav rms %.3e (signal rms %.3e), oracle %.0f s"
          % (err, rms, ref_wav.pow(2).mean().sqrt().item(), t_oracle))
    assert err <= 1e-3, err                      # north_star gate on the denoised mel
    assert rms <= 1e-4, rms                      # north_star gate on the waveform


def test_cfg2_plms_interval40_x_862():
    _threads()
    gd, sd = _model()
    cond, x0, _ = _inputs(1, 862, 0, seed=11)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    with torch.no_grad():
        ref = O.sample(sd, sched, cond, x0, 1000, None, pndm_speedup=40)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), 1000, 40).cpu()
    rng = max(1.0, ref.abs().max().item())       # PLMS has no clamp (diffusion.py:166-198): bound relative to the range
    err = (xf - ref).abs().max().item()
    print("cfg2 plms d40 x 862: max-abs %.3e, range %.3e" % (err, rng))
    assert err / rng <= 2e-5, (err, rng)
    mel = O.mel_from_x(xf, SPEC_MIN, SPEC_MAX); ref_mel = O.mel_from_x(ref, SPEC_MIN, SPEC_MAX)
    assert (mel - ref_mel).abs().max().item() <= 1e-3 * max(1.0, rng)


def test_cfg4_plms_interval20_x_43():
    """The flask chunk (0.5 s = 43 frames, 51 evals): one frame tile, the short-clip kernel selection."""
    _threads()
    gd, sd = _model()
    cond, x0, _ = _inputs(1, 43, 0, seed=12)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    with torch.no_grad():
        ref = O.sample(sd, sched, cond, x0, 1000, None, pndm_speedup=20)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), 1000, 20).cpu()
    rng = max(1.0, ref.abs().max().item())
    err = (xf - ref).abs().max().item()
    print("cfg4 plms d20 x 43: max-abs %.3e, range %.3e" % (err, rng))
    assert err / rng <= 2e-5, (err, rng)


def test_cfg3_ragged_8_x_689():
    """8 slices of ~8 s (689 frames +- 25 %) as one ragged batch; every item must equal its own B=1 oracle run."""
    _threads()
    steps = 24
    lens = [689, 861, 517, 700, 640, 689, 803, 575]
    T = max(lens)
    gd, sd = _model(K_step=steps)
    cond, x0, noise = _inputs(len(lens), T, steps, seed=21)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    worst = 0.0
    for b, n in enumerate(lens):
        with torch.no_grad():
            ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
        worst = max(worst, (xf[b:b + 1, :, :, :n] - ref).abs().max().item())
    print("cfg3 ragged 8x~689, %d steps: worst max-abs %.3e" % (steps, worst))
    assert worst <= 2e-4, worst


@pytest.mark.parametrize("math_mode,tol", [("fp32", 1e-5), ("tc3f16", 5e-5)])
@pytest.mark.parametrize("B,T", [(1, 1000), (4, 862)])
def test_one_residual_block_per_dilation(math_mode, tol, B, T):
    """SURVEY.md section 7 minimum slice: a net whose LAST layer has dilation 1, 2, 4, 8 (L = 1..4, cycle 4); the
    error of eps is dominated by the deepest block, so each dilation's kernels are pinned at the BASELINE widths."""
    _threads()
    g = torch.Generator().manual_seed(100 + T)
    for L in (1, 2, 3, 4):
        import diffsvc_b200 as D
        _hp(residual_layers=L, dilation_cycle_length=4)
        sd = O.synth_diffnet_weights(L=L, seed=50 + L)
        dn = D.DiffNet(128, math_mode=math_mode)
        dn.load_state_dict(sd, strict=True)
        dn = dn.to(DEV)
        spec = torch.randn(B, 1, 128, T, generator=g)
        cond = torch.randn(B, 256, T, generator=g) * 0.5
        t = torch.full((B,), 417, dtype=torch.long)
        with torch.no_grad():
            ref = O.diffnet_forward(sd, spec, t, cond)
        out = dn(spec.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
        err = (out - ref).abs().max().item()
        assert err <= tol, (math_mode, L, 2 ** (L - 1), err)


def test_cosine_schedule_chain():
    """schedule_type='cosine' (diffusion.py:48-58,:74-77) through the product constructor vs the oracle's schedule."""
    _threads()
    steps, T = 40, 200
    gd, sd = _model(K_step=steps, schedule_type="cosine")
    sched = O.make_schedule(O.cosine_beta_schedule(1000))
    assert abs(float(gd.betas[-1]) - float(sched["betas"][-1])) <= 1e-7 and float(gd.betas[-1]) > 0.5
    cond, x0, noise = _inputs(1, T, steps, seed=33)
    with torch.no_grad():
        ref = O.sample(sd, sched, cond, x0, steps, noise)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV)).cpu()
    err = (xf - ref).abs().max().item()
    assert err <= 3e-4, err
