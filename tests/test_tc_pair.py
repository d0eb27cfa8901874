"""The CTA-pair main loop (csrc/tc_pair.cuh: tcgen05 cta_group::2, M = 256, each SM stages half of the weight tile) --
the default for the 64- and 128-wide tile classes of the 3-pass mode -- against the CPU oracle and against the
single-CTA kernels (DSVC_TC_PAIR=0).  Same three partial products per channel, another summation order for half of the
channels: equal to fp32 rounding, not bit-identical."""
import pytest
import torch

from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(monkeypatch, pair, K_step=1000, bn=None):
    import diffsvc_b200 as D
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
    monkeypatch.setenv("DSVC_TC_PAIR", "1" if pair else "0")
    monkeypatch.setenv("DSVC_SPLITK", "0")
    if bn:
        monkeypatch.setenv("DSVC_TC_BN", str(bn))
    else:
        monkeypatch.delenv("DSVC_TC_BN", raising=False)
    sd = O.synth_diffnet_weights()
    dn = D.DiffNet(128, math_mode="tc3f16")
    dn.load_state_dict(sd, strict=True)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=K_step, loss_type="l2", spec_min=[-5.0], spec_max=[0.0])
    return gd.to(DEV).eval(), sd


def _inputs(B, T, steps, seed=7):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 256, T, generator=g) * 0.5, torch.randn(B, 1, 128, T, generator=g),
            torch.randn(steps, B, 1, 128, T, generator=g))


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("B,T", [(1, 862), (1, 43), (1, 129), (2, 300), (1, 1)])
def test_eval_vs_oracle_and_single_cta(monkeypatch, B, T, bn):
    """One DiffNet evaluation: odd and even numbers of frame tiles (the last pair half empty), both tile widths."""
    cond, x0, _ = _inputs(B, T, 1)
    t = torch.full((B,), 417, dtype=torch.long)
    gd, sd = _model(monkeypatch, True, bn=bn)
    out = gd.denoise_fn(x0.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
    with torch.no_grad():
        ref = O.diffnet_forward(sd, x0, t, cond)
    err = (out - ref).abs().max().item()
    assert torch.isfinite(out).all() and err <= 1e-4, (B, T, bn, err)
    gd1, _ = _model(monkeypatch, False, bn=bn)
    one = gd1.denoise_fn(x0.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
    assert (out - one).abs().max().item() <= 2e-5 * max(1.0, one.abs().max().item())
    assert not torch.equal(out, one)                    # the pair kernels did run


def test_ddpm_chain_ragged_and_deterministic(monkeypatch):
    steps, lens = 12, [300, 129, 128, 5]
    cond, x0, noise = _inputs(len(lens), max(lens), steps, seed=21)
    gd, sd = _model(monkeypatch, True)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    run = lambda: gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    xf = run()
    assert torch.equal(xf, run())
    for b, n in enumerate(lens):
        with torch.no_grad():
            ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
        assert (xf[b:b + 1, :, :, :n] - ref).abs().max().item() <= 2e-4, b
    # an item's result does not depend on its batch (same tile class -> bit-identical)
    one = gd.sample(x0[1:2, :, :, :129].contiguous().to(DEV), cond[1:2, :, :129].contiguous().to(DEV), steps, None,
                    noise[:, 1:2, :, :, :129].contiguous().to(DEV)).cpu()
    assert torch.equal(one, xf[1:2, :, :, :129])


@pytest.mark.parametrize("bn", [None, 128, 256])
def test_vocoder_on_pair_kernels(monkeypatch, bn):
    """NSF-HiFiGAN ResBlock convs (k = 3 / 7 / 11 taps, dilations 1 / 3 / 5) through the same main loop, in every tile
    class (a 10 s clip takes the 128-wide class in its first stages, batches the 256-wide one)."""
    if bn:
        monkeypatch.setenv("DSVC_TC_BN", str(bn))
    else:
        monkeypatch.delenv("DSVC_TC_BN", raising=False)
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear(); hparams.update(DEFAULTS_44K)
    sd = O.synth_nsf_weights(O.NSF_H_44K)
    B, T = 2, 33
    g = torch.Generator().manual_seed(3)
    mel = torch.randn(B, T, 128, generator=g) * 0.8 - 2.0
    f0 = O.synth_f0(B, T)
    rand_ini = torch.rand(B, 9, generator=g)
    noise = torch.randn(B, T * 512, 9, generator=g)
    with torch.no_grad():
        ref = O.spec2wav(sd, O.NSF_H_44K, mel, f0, rand_ini, noise)
    outs = {}
    for pair in (True, False):
        monkeypatch.setenv("DSVC_TC_PAIR", "1" if pair else "0")
        voc = NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device=DEV)
        outs[pair] = voc.spec2wav_torch(mel.to(DEV), f0=f0.to(DEV), rand_ini=rand_ini, sine_noise=noise).cpu()
        d = outs[pair] - ref
        assert d.pow(2).mean().sqrt().item() <= 1e-4 and d.abs().max().item() <= 5e-4, pair
    assert (outs[True] - outs[False]).abs().max().item() <= 5e-5


@pytest.mark.parametrize("pack", ["0", "1"])
def test_ragged_tile_table_follows_the_lengths(monkeypatch, pack):
    """DSVC_PACK=0: ragged batches run from a table of the frame tiles that hold a valid frame (dead slots exit).  One
    handle, one (B, Tmax), three different length sets back to back -- the captured step graph is reused, only the table
    changes -- then full-length again (dense grid).  DSVC_PACK=1 (default): the same calls on the packed frame axis
    (row map rebuilt per length set; the graph is re-captured when the packed length changes its 256-frame tile count)."""
    monkeypatch.setenv("DSVC_PACK", pack)
    steps, B, T = 5, 3, 400
    cond, x0, noise = _inputs(B, T, steps, seed=77)
    gd, sd = _model(monkeypatch, True, K_step=steps)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    for lens in ([400, 130, 7], [1, 400, 256], [129, 128, 0], None):
        xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
        for b, n in enumerate(lens or [T] * B):
            if n == 0:
                continue
            with torch.no_grad():
                ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
            err = (xf[b:b + 1, :, :, :n] - ref).abs().max().item()
            assert err <= 1e-4, (lens, b, err)


def test_vocoder_narrow_stages_on_tensor_cores(monkeypatch):
    """The 32- / 16-channel ResBlock stages on the weights-stationary narrow kernel (csrc/tc_narrow.cuh: 64- / 32-byte
    operand rows, SWIZZLE_64B / 32B, every tap fed from one 192-row activation window at a row offset): on
    (DSVC_NSF_NARROW=1), off (=0: FFMA GEMM) and the default (on), all against the CPU oracle."""
    from diffsvc_b200.vocoders.nsf_hifigan import NsfHifiGAN
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear(); hparams.update(DEFAULTS_44K)
    monkeypatch.delenv("DSVC_TC_BN", raising=False)
    monkeypatch.setenv("DSVC_TC_PAIR", "1")
    sd = O.synth_nsf_weights(O.NSF_H_44K)
    B, T = 2, 41
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(B, T, 128, generator=g) * 0.8 - 2.0
    f0 = O.synth_f0(B, T)
    rand_ini = torch.rand(B, 9, generator=g)
    noise = torch.randn(B, T * 512, 9, generator=g)
    with torch.no_grad():
        ref = O.spec2wav(sd, O.NSF_H_44K, mel, f0, rand_ini, noise)
    outs = {}
    for narrow in ("1", "0", None):
        if narrow is None:
            monkeypatch.delenv("DSVC_NSF_NARROW", raising=False)
        else:
            monkeypatch.setenv("DSVC_NSF_NARROW", narrow)
        voc = NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device=DEV)
        outs[narrow] = voc.spec2wav_torch(mel.to(DEV), f0=f0.to(DEV), rand_ini=rand_ini, sine_noise=noise).cpu()
        d = outs[narrow] - ref
        assert d.pow(2).mean().sqrt().item() <= 1e-4 and d.abs().max().item() <= 5e-4, narrow
    assert not torch.equal(outs["1"], outs["0"])          # the narrow-row kernels did run
    assert torch.equal(outs[None], outs["1"])             # the default
