"""PitchExtractor (mel -> f0, 24 kHz models): oracle vs the golden dumped from the
reference's module (CPU), the CUDA path through the C-ABI vs both (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import diffsvc_oracle as O  # noqa: E402

import diffsvc_b200 as D  # noqa: E402
from diffsvc_b200 import _lib  # noqa: E402
from diffsvc_b200.pe import PitchExtractor  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "pe_small.npz"))
SD = {k[3:]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith("sd/")}


@pytest.fixture
def small_hparams():
    keys = ("hidden_size", "predictor_hidden", "predictor_kernel", "ffn_padding", "pitch_norm", "pitch_type", "use_uv",
            "f0_mean", "f0_std")
    old = {k: D.hparams.get(k) for k in keys}
    D.hparams.update(hidden_size=64, predictor_hidden=-1, predictor_kernel=5, ffn_padding="SAME", pitch_norm="log",
                     pitch_type="frame", use_uv=False)
    yield D.hparams
    for k, v in old.items():
        if v is None:
            D.hparams.pop(k, None)
        else:
            D.hparams[k] = v


def test_oracle_matches_reference_golden():
    pred, f0 = O.pitch_extractor(SD, torch.from_numpy(GOLD["mel"]))
    assert np.abs(pred.numpy() - GOLD["pitch_pred"]).max() <= 1e-5
    assert np.abs(f0.numpy() - GOLD["f0_denorm_pred"]).max() <= 1e-3 * 1     # Hz (values ~ 200)
    assert (f0.numpy()[1, 37:] == 0).all() and (GOLD["f0_denorm_pred"][1, 37:] == 0).all()


def test_state_dict_matches_reference_keys(small_hparams):
    m = PitchExtractor(n_mel_bins=80, conv_layers=2)
    ours = m.state_dict()
    assert set(ours) == set(SD), set(ours) ^ set(SD)
    for k in SD:
        assert tuple(ours[k].shape) == tuple(SD[k].shape), k
    m.load_state_dict(SD, strict=True)


def test_no_cpu_path(small_hparams):
    m = PitchExtractor()
    with pytest.raises(_lib.DsvcError):
        m(torch.zeros(1, 4, 80))


def _run(m, mel):
    ret = m(mel.cuda())
    return ret["pitch_pred"].cpu().numpy(), ret["f0_denorm_pred"].cpu().numpy()


@pytest.mark.gpu
def test_kernels_match_reference_golden(small_hparams):
    m = PitchExtractor(80, 2)
    m.load_state_dict(SD, strict=True)
    m = m.cuda()
    pred, f0 = _run(m, torch.from_numpy(GOLD["mel"]))
    err = np.abs(pred - GOLD["pitch_pred"]).max()
    rel = np.abs(f0 - GOLD["f0_denorm_pred"]).max() / np.abs(GOLD["f0_denorm_pred"]).max()
    print("pitch extractor: pitch_pred max-abs %.2e, f0 rel %.2e" % (err, rel))
    assert err <= 2e-4 and rel <= 2e-4
    assert (f0[1, 37:] == 0).all()
    # same call again (cached handle), then after a weight change (rebuilt handle)
    pred2, _ = _run(m, torch.from_numpy(GOLD["mel"]))
    assert np.array_equal(pred, pred2)
    with torch.no_grad():
        m.pitch_predictor.linear.bias.add_(1.0)
    pred3, _ = _run(m, torch.from_numpy(GOLD["mel"]))
    assert np.abs(pred3 - pred - 1.0).max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["standard_uv", "causal", "no_encoder", "long"])
def test_kernels_match_oracle_variants(small_hparams, variant):
    kw = dict(conv_layers=2, pad_same=True, pitch_norm="log", use_uv=False, f0_mean=0.0, f0_std=1.0)
    T = 77
    if variant == "standard_uv":
        kw.update(pitch_norm="standard", use_uv=True, f0_mean=200.0, f0_std=30.0)
    if variant == "causal":
        kw.update(pad_same=False)
    if variant == "no_encoder":
        kw.update(conv_layers=0)
    if variant == "long":
        T = 4200                                       # beyond the 4096-row position table
    small_hparams.update(pitch_norm=kw["pitch_norm"], use_uv=kw["use_uv"], f0_mean=kw["f0_mean"], f0_std=kw["f0_std"],
                         ffn_padding="SAME" if kw["pad_same"] else "LEFT")
    torch.manual_seed(3)
    m = PitchExtractor(80, kw["conv_layers"])
    with torch.no_grad():
        for n, b in m.named_buffers():
            if n.endswith("running_var"):
                b.copy_(0.5 + torch.rand_like(b))
            if n.endswith("running_mean"):
                b.copy_(0.1 * torch.randn_like(b))
        m.pitch_predictor.linear.bias.add_(torch.tensor([0.3 if variant == "standard_uv" else 7.0, 0.0]))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    mel = torch.randn(2, T, 80) - 3.0
    mel[1, T - 9:] = 0
    want_pred, want_f0 = O.pitch_extractor(sd, mel, **kw)
    pred, f0 = _run(m.cuda(), mel)
    assert np.abs(pred - want_pred.numpy()).max() <= 2e-4
    assert np.abs(f0 - want_f0.numpy()).max() <= 2e-4 * max(1.0, float(want_f0.abs().max()))
    if variant == "standard_uv":
        assert (f0 == 0).any() and (f0 != 0).any()


@pytest.mark.gpu
def test_empty_inputs(small_hparams):
    m = PitchExtractor(80, 2).cuda()
    ret = m(torch.zeros(0, 5, 80).cuda())
    assert ret["pitch_pred"].shape == (0, 5, 2) and ret["f0_denorm_pred"].shape == (0, 5)
    ret = m(torch.zeros(2, 0, 80).cuda())
    assert ret["f0_denorm_pred"].shape == (2, 0)
