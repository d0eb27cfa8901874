"""`NsfHifiGAN` -- drop-in for network/vocoders/nsf_hifigan.py:9-92 (registered under the same names)."""
import os

import torch

from ..hparams import hparams
from .base_vocoder import BaseVocoder, register_vocoder
from .nsf_models import Generator, load_model  # noqa: F401  (re-exported like the reference's imports)
from .nvstft import STFT, load_wav_to_torch  # noqa: F401

_CHECKS = (("sampling_rate", "audio_sample_rate"), ("num_mels", "audio_num_mel_bins"), ("n_fft", "fft_size"),
           ("win_size", "win_size"), ("hop_size", "hop_size"), ("fmin", "fmin"), ("fmax", "fmax"))


@register_vocoder
class NsfHifiGAN(BaseVocoder):
    def __init__(self, device=None):
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = device
        model_path = hparams["vocoder_ckpt"]
        if os.path.exists(model_path):
            print("| Load HifiGAN: ", model_path)
            self.model, self.h = load_model(model_path, device=self.device)
        else:
            print("Error: HifiGAN model file is not found!")

    @classmethod
    def from_state_dict(cls, h, state_dict, device="cuda"):
        """Build from in-memory (synthetic) weights: the reference ships no checkpoints."""
        self = cls.__new__(cls)
        self.device = device
        self.model = Generator(h, state_dict, device=device)
        self.h = self.model.h
        return self

    def _check(self):
        for hk, pk in _CHECKS:
            if hk in self.h and pk in hparams and self.h[hk] != hparams[pk]:
                print("Mismatch parameters: hparams['%s']=" % pk, hparams[pk], "!=", self.h[hk], "(vocoder)")

    def spec2wav_torch(self, mel, **kwargs):  # mel: [B, T, bins]
        self._check()
        with torch.no_grad():
            f0 = kwargs.get("f0")  # [B, T]
            if f0 is None or not hparams.get("use_nsf"):
                raise NotImplementedError("the NSF generator needs f0 (use_nsf)")
            extra = {k: kwargs[k] for k in ("rand_ini", "sine_noise", "seed") if k in kwargs}
            # c = 2.30259 * mel (log10 -> ln) is applied on load inside the kernel
            return self.model.forward_mel(mel.to(self.device), f0.to(self.device), 2.30259, **extra).view(-1)

    def spec2wav(self, mel, **kwargs):
        self._check()
        with torch.no_grad():
            c = torch.FloatTensor(mel).unsqueeze(0).to(self.device)          # [1, T, bins]
            f0 = kwargs.get("f0")
            if f0 is None or not hparams.get("use_nsf"):
                raise NotImplementedError("the NSF generator needs f0 (use_nsf)")
            f0 = torch.FloatTensor(f0[None, :]).to(self.device)
            extra = {k: kwargs[k] for k in ("rand_ini", "sine_noise", "seed") if k in kwargs}
            y = self.model.forward_mel(c, f0, 2.30259, **extra).view(-1)
        return y.cpu().numpy()

    @staticmethod
    def wav2spec(inp_path, device=None):
        """network/vocoders/nsf_hifigan.py:75-92: (wav np[L], log10-mel np[T, M]); the analysis is one kernel."""
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        stft = _stft_for(hparams)
        with torch.no_grad():
            wav_torch, _ = load_wav_to_torch(inp_path, target_sr=stft.target_sr)
            # log mel to log10 mel (0.434294) and the [T, M] transpose happen inside the kernel
            mel_torch = stft.get_mel(wav_torch.unsqueeze(0).to(device), out_scale=0.434294, transpose=False).squeeze(0)
            return wav_torch.cpu().numpy(), mel_torch.cpu().numpy()

    def spec2wav_device(self, mel, f0, **kwargs):
        """mel [T, M] / f0 [T] CUDA tensors -> CUDA waveform [T*hop]; `spec2wav` without the host round trip
        (used by diffsvc_b200.infer_glue.after_infer)."""
        return self.spec2wav_torch(mel.unsqueeze(0), f0=f0.unsqueeze(0), **kwargs)


_STFTS = {}


def _stft_for(hp):
    key = tuple(hp[k] for k in ("audio_sample_rate", "audio_num_mel_bins", "fft_size", "win_size", "hop_size",
                                "fmin", "fmax"))
    if key not in _STFTS:
        _STFTS[key] = STFT(*key)
    return _STFTS[key]
