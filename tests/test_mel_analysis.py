"""Mel analysis (`wav2spec`) and the device-side `after_infer` glue.

CPU: the oracle against the golden vectors dumped from the reference's STFT.get_mel, the librosa mel basis
restatements (oracle and product) against each other and against torchaudio, the host-only entry points.
GPU: the kernels through the C-ABI against the oracle / goldens.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import diffsvc_oracle as O  # noqa: E402

import diffsvc_b200 as D  # noqa: E402
from diffsvc_b200 import _lib, infer_glue  # noqa: E402
from diffsvc_b200.vocoders import nvstft  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "mel_small.npz"))
TAGS = ["a", "b", "c"]


def _cfg(tag):
    sr, n_mels, n_fft, win, hop, fmin, fmax = [int(v) for v in GOLD[tag + "/cfg"]]
    return dict(sr=sr, n_mels=n_mels, n_fft=n_fft, win=win, hop=hop, fmin=fmin, fmax=fmax)


# ---------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden(tag):
    c = _cfg(tag)
    wav = torch.from_numpy(GOLD[tag + "/wav"]).unsqueeze(0)
    mel = O.mel_analysis(wav, c["n_fft"], c["win"], c["hop"], GOLD[tag + "/basis"]).squeeze(0).numpy()
    assert mel.shape == GOLD[tag + "/mel_ln"].shape
    assert np.abs(mel - GOLD[tag + "/mel_ln"]).max() <= 1e-5          # same torch build, same ops
    # the fp64 arbiter agrees with the reference's fp32 run up to fp32 FFT noise
    mel64 = O.mel_analysis(wav, c["n_fft"], c["win"], c["hop"], GOLD[tag + "/basis"], dtype=torch.float64).squeeze(0).numpy()
    assert np.abs(mel64 - GOLD[tag + "/mel_ln"]).max() <= 5e-3


@pytest.mark.parametrize("tag", TAGS)
def test_mel_basis_restatements_agree(tag):
    c = _cfg(tag)
    ours = nvstft.slaney_mel_basis(c["sr"], c["n_fft"], c["n_mels"], c["fmin"], c["fmax"])
    orc = O.slaney_mel_basis(c["sr"], c["n_fft"], c["n_mels"], c["fmin"], c["fmax"])
    assert ours.dtype == np.float32 and ours.shape == (c["n_mels"], c["n_fft"] // 2 + 1)
    assert np.abs(ours - orc).max() <= 1e-9
    assert np.array_equal(orc, GOLD[tag + "/basis"])
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(c["n_fft"] // 2 + 1, float(c["fmin"]), float(c["fmax"]), c["n_mels"], c["sr"],
                                       norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(ours - fb).max() <= 2e-6 * max(1.0, float(fb.max()) / 0.03)   # independent fp32 implementation


def test_band_ranges():
    basis = nvstft.slaney_mel_basis(44100, 2048, 128, 40, 16000)
    lo, hi = nvstft.band_ranges(basis)
    for m in range(128):
        assert (basis[m, :lo[m]] == 0).all() and (basis[m, hi[m]:] == 0).all()
        assert basis[m, lo[m]] != 0 and basis[m, hi[m] - 1] != 0
    z = np.zeros((2, 5), np.float32)
    lo, hi = nvstft.band_ranges(z)
    assert (lo == 0).all() and (hi == 0).all()


def test_mel_frames_host_entry_point():
    lib = _lib.load()
    for tag in TAGS:
        c = _cfg(tag)
        cfg = _lib.MelConfig(c["n_fft"], c["hop"], c["n_mels"], 1e-5, 1.0)
        assert lib.dsvc_mel_frames(cfg, len(GOLD[tag + "/wav"])) == GOLD[tag + "/mel_ln"].shape[1]
    cfg = _lib.MelConfig(2048, 512, 128, 1e-5, 1.0)
    assert lib.dsvc_mel_frames(cfg, 768) == -1           # reflect padding needs more than (n_fft-hop)/2 samples
    assert lib.dsvc_mel_frames(cfg, 769) == 1
    assert lib.dsvc_mel_frames(cfg, 44100 * 10) == 861    # floor(n / hop) for even n_fft - hop


def test_mel_analysis_has_no_cpu_path():
    stft = nvstft.STFT(44100, 128, 2048, 2048, 512, 40, 16000)
    with pytest.raises(_lib.DsvcError):
        stft.get_mel(torch.zeros(1, 4096))
    with pytest.raises(_lib.DsvcError):
        infer_glue.compact_frames(torch.zeros(4, 8), None, -6.0, 1.5)


def test_after_infer_defers_to_reference_for_host_tensors():
    calls = []

    class Svc:
        vocoder = object()

        def after_infer(self, prediction, singer, in_path):
            calls.append(prediction)
            return "reference"

    infer_glue.patch(Svc)
    infer_glue.patch(Svc)                                   # idempotent
    assert Svc._dsvc_reference_after_infer is not Svc.after_infer
    out = Svc().after_infer({"outputs": torch.zeros(1, 4, 8), "mels": torch.zeros(1, 4, 8), "f0_pred": torch.zeros(1, 4)}, False, "x.wav")
    assert out == "reference" and len(calls) == 1


def test_oracle_after_infer_frames():
    mel = np.random.default_rng(0).standard_normal((6, 4)).astype(np.float32) * 4
    mel[[1, 4]] = 0
    f0 = np.arange(6, dtype=np.float32)
    m, f = O.after_infer_frames(mel, f0, -6.0, 1.5)
    assert m.shape == (4, 4) and list(f) == [0, 2, 3, 5] and m.max() <= 1.5 and m.min() >= -6.0


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_mel_kernel_vs_oracle_and_golden(tag):
    c = _cfg(tag)
    wav = torch.from_numpy(GOLD[tag + "/wav"]).unsqueeze(0)
    stft = nvstft.STFT(c["sr"], c["n_mels"], c["n_fft"], c["win"], c["hop"], c["fmin"], c["fmax"])
    got = stft.get_mel(wav.cuda()).squeeze(0).cpu().numpy()                 # [n_mels, T] natural log
    gold = GOLD[tag + "/mel_ln"]
    assert got.shape == gold.shape
    o64 = O.mel_analysis(wav, c["n_fft"], c["win"], c["hop"], GOLD[tag + "/basis"], dtype=torch.float64).squeeze(0).numpy()
    err_ours = np.abs(got - o64).max()
    err_ref = np.abs(gold - o64).max()
    print("mel %s: |kernel - fp64| %.2e   |reference fp32 - fp64| %.2e   |kernel - reference| %.2e"
          % (tag, err_ours, err_ref, np.abs(got - gold).max()))
    assert err_ours <= 2e-5                      # fp64 DFT inside: the only fp32 steps are magnitude, mel dot, log
    assert np.abs(got - gold).max() <= err_ref + 2e-5   # and never further from the reference than its own FFT noise


@pytest.mark.gpu
def test_wav2spec_file_round_trip(tmp_path):
    from scipy.io import wavfile
    c = _cfg("a")
    pcm = np.round(GOLD["a/wav"] * 32767).astype(np.int16)
    path = str(tmp_path / "clip.wav")
    wavfile.write(path, c["sr"], pcm)
    D.hparams.update(audio_sample_rate=c["sr"], audio_num_mel_bins=c["n_mels"], fft_size=c["n_fft"], win_size=c["win"],
                     hop_size=c["hop"], fmin=c["fmin"], fmax=c["fmax"])
    wav, mel = D.NsfHifiGAN.wav2spec(path)
    assert wav.dtype == np.float32 and mel.dtype == np.float32
    assert np.array_equal(wav, pcm.astype(np.float32) / 32768)            # nvSTFT.py:30-36 normalisation
    ref = O.wav2spec(torch.from_numpy(wav), c["n_fft"], c["win"], c["hop"], GOLD["a/basis"], dtype=torch.float64).numpy()
    assert mel.shape == ref.shape == (len(wav) // c["hop"], c["n_mels"])
    assert np.abs(mel - ref).max() <= 2e-5


@pytest.mark.gpu
def test_mel_kernel_long_and_batched():
    g = torch.Generator().manual_seed(3)
    wav = (torch.rand(2, 44100 * 3, generator=g) * 2 - 1) * 0.5
    stft = nvstft.STFT(44100, 128, 2048, 2048, 512, 40, 16000)
    got = stft.get_mel(wav.cuda()).cpu()
    basis = O.slaney_mel_basis(44100, 2048, 128, 40, 16000)
    ref = O.mel_analysis(wav, 2048, 2048, 512, basis, dtype=torch.float64)
    assert got.shape == ref.shape == (2, 128, 258)
    assert (got - ref).abs().max().item() <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("T,M,zero_every", [(862, 128, 7), (3000, 128, 3), (5, 80, 1), (1, 128, 0), (1500, 16, 0)])
def test_compact_frames_bit_exact(T, M, zero_every):
    g = np.random.default_rng(T)
    mel = (g.standard_normal((T, M)) * 3 - 2).astype(np.float32)
    if zero_every:
        mel[::zero_every] = 0
    mel[T // 2, : M // 2] = 0                      # a partly-zero row is kept
    f0 = g.uniform(80, 800, T).astype(np.float32)
    want_mel, want_f0 = O.after_infer_frames(mel, f0, -6.0, 1.5)
    got_mel, got_f0 = infer_glue.compact_frames(torch.from_numpy(mel).cuda(), torch.from_numpy(f0).cuda(), -6.0, 1.5)
    assert np.array_equal(got_mel.cpu().numpy(), want_mel)
    assert np.array_equal(got_f0.cpu().numpy(), want_f0)
    only_mel, none = infer_glue.compact_frames(torch.from_numpy(mel).cuda(), None, -6.0, 1.5)
    assert none is None and np.array_equal(only_mel.cpu().numpy(), want_mel)


@pytest.mark.gpu
def test_after_infer_on_device_matches_host_path():
    """Svc.after_infer semantics (infer_tool.py:172-200) with CUDA tensors in the prediction dict."""
    h = dict(O.NSF_H_44K, upsample_initial_channel=128)
    sd = O.synth_nsf_weights(h, seed=5)
    D.hparams.update(use_nsf=True, mel_vmin=-6.0, mel_vmax=1.5, audio_sample_rate=44100, audio_num_mel_bins=128,
                     fft_size=2048, win_size=2048, hop_size=512, fmin=40, fmax=16000)
    voc = D.NsfHifiGAN.from_state_dict(h, sd)

    class Voc:                                               # pin the vocoder's random draws for the comparison
        def spec2wav_device(self, mel, f0):
            return voc.spec2wav_device(mel, f0, seed=5)

    class Svc:
        vocoder = Voc()

        def after_infer(self, prediction, singer, in_path):
            raise AssertionError("the device path should have handled this")

    infer_glue.patch(Svc)
    T = 40
    g = np.random.default_rng(1)
    mel = (g.standard_normal((1, T, 128)) * 2 - 3).astype(np.float32)
    mel[0, 30:] = 0                                          # padding frames
    mels_gt = mel.copy()
    f0 = g.uniform(100, 400, (1, T)).astype(np.float32)
    pred = {"outputs": torch.from_numpy(mel).cuda(), "mels": torch.from_numpy(mels_gt), "f0_gt": torch.from_numpy(f0 * 2),
            "f0_pred": torch.from_numpy(f0).cuda(), "mel2ph_pred": None}
    f0_gt, f0_pred, wav = Svc().after_infer(pred, False, "x.wav")
    want_mel, want_f0 = O.after_infer_frames(mel[0], f0[0], -6.0, 1.5)
    assert np.array_equal(f0_pred, want_f0) and np.array_equal(f0_gt, (f0 * 2)[0][:30])
    assert isinstance(pred["outputs"], np.ndarray)           # the dict is converted like the reference does
    want_wav = voc.spec2wav(want_mel, f0=want_f0, seed=5)
    assert wav.shape == (30 * 512,) and np.array_equal(wav, want_wav)
