"""The step kernel (csrc/tc_step.cuh, opt-in DSVC_STEP=1): the 2L+3 contractions of one DiffNet evaluation as phases of ONE
launch of resident CTA pairs (warp-specialised: TMA producer / MMA issuer / 16 epilogue warps, two accumulators in TMEM,
per-frame-tile dependency counters instead of kernel boundaries).  Same tile math and the same epilogue functors as the
per-layer pair kernels: bit-identical to them at the same tile width, and within the parity gate against the CPU oracle.
Measured slower than the per-layer kernels, so it stays opt-in -- but it stays tested."""
import pytest
import torch

import diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(monkeypatch, step, K_step=1000):
    monkeypatch.setenv("DSVC_STEP", "1" if step else "0")
    monkeypatch.delenv("DSVC_TC_BN", raising=False)
    monkeypatch.delenv("DSVC_SPLITK", raising=False)
    hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
    sd = O.synth_diffnet_weights()
    dn = D.DiffNet(128, math_mode="tc3f16")
    dn.load_state_dict(sd, strict=True)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=K_step, loss_type="l2", spec_min=[-5.0], spec_max=[0.0]).to(DEV).eval()
    return gd, sd


def _inputs(B, T, steps, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 256, T, generator=g) * 0.5, torch.randn(B, 1, 128, T, generator=g),
            torch.randn(steps, B, 1, 128, T, generator=g))


@pytest.mark.parametrize("B,T,lens", [(1, 300, None), (1, 43, None), (3, 150, [150, 97, 33])])
def test_step_kernel_bit_identical_to_per_layer_kernels(monkeypatch, B, T, lens):
    """One clip / a small packed batch: 64-wide slots, one per CTA pair = the per-layer kernels' tile class."""
    steps = 6
    cond, x0, noise = _inputs(B, T, steps, seed=11)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    out = {}
    for step in (True, False):
        gd, sd = _model(monkeypatch, step)
        out[step] = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
        again = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
        assert torch.equal(out[step], again)                       # deterministic, handle reusable
        plms = gd.sample(x0.to(DEV), cond.to(DEV), 100, 20, None, lengths=lens).cpu()
        out[(step, "plms")] = plms
        t = torch.full((B,), 37, dtype=torch.long)
        out[(step, "eval")] = gd.denoise_fn(x0.to(DEV), t.to(DEV), cond.to(DEV)).cpu() if lens is None else None
    assert torch.equal(out[True], out[False])
    assert torch.equal(out[(True, "plms")], out[(False, "plms")])
    if lens is None:
        assert torch.equal(out[(True, "eval")], out[(False, "eval")])
    for b in range(B):
        n = lens[b] if lens else T
        with torch.no_grad():
            ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
        assert (out[True][b:b + 1, :, :, :n] - ref).abs().max().item() <= 1e-4, b


def test_step_kernel_multi_slot_batch(monkeypatch):
    """A batch too large for one slot per pair: 128-wide slots, two per CTA pair, the pair's producer / issuer run ahead
    of its epilogue across slots and phases.  Another tile class than the per-layer kernels pick (256-wide) -> equal to
    fp32 rounding; every item within the gate against its own B = 1 oracle run."""
    steps, lens = 4, [700, 650, 512, 1, 0, 300, 699, 257]
    B, T = len(lens), max(lens)
    cond, x0, noise = _inputs(B, T, steps, seed=12)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    gd, sd = _model(monkeypatch, True)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    assert torch.equal(xf, gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu())
    gd0, _ = _model(monkeypatch, False)
    x0f = gd0.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    for b in (0, 3, 5, 7):
        n = lens[b]
        assert (xf[b, :, :, :n] - x0f[b, :, :, :n]).abs().max().item() <= 2e-5
        with torch.no_grad():
            ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
        assert (xf[b:b + 1, :, :, :n] - ref).abs().max().item() <= 1e-4, b
