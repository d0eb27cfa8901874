"""`HifiGAN` / `HifiGanGenerator` -- drop-in for the 24 kHz vocoder: network/vocoders/hifigan.py:17-76 and
modules/hifigan/hifigan.py:104-169.  The generator is the same network as the NSF one (identical
SineGen / SourceModuleHnNSF, modules/parallel_wavegan/models/source.py:484); it runs on the same
libdsvc kernels with `mel_scale = 1` (this wrapper feeds the mel unscaled, hifigan.py:65) and an
optional harmonic source (`use_pitch_embed`, `f0=None` skips it)."""
import glob
import json
import os
import re

import torch

from ..hparams import hparams
from .base_vocoder import BaseVocoder, register_vocoder
from .nsf_models import AttrDict, Generator


class HifiGanGenerator(Generator):
    """modules/hifigan/hifigan.py:104: config keys `audio_sample_rate`, fixed 80 mel bins (:118)."""

    def __init__(self, h, c_out=1, state_dict=None, device="cuda"):
        h = AttrDict(dict(h))
        h.setdefault("num_mels", 80)
        h.setdefault("sampling_rate", h.get("audio_sample_rate", 24000))
        super().__init__(h, state_dict, device=device)


def load_model(config_path, file_path, device=None):
    """network/vocoders/hifigan.py:17-39 (.ckpt branch; the reference's .pth branch pickles a whole module)."""
    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    ext = os.path.splitext(file_path)[-1]
    if ext != ".ckpt":
        raise NotImplementedError("only .ckpt generator checkpoints are supported")
    ckpt = torch.load(file_path, map_location="cpu")
    if ".yaml" in config_path:
        import yaml
        with open(config_path, encoding="utf-8") as f:
            config = yaml.safe_load(f)
        state = ckpt["state_dict"]["model_gen"]
    else:
        config = json.load(open(config_path, "r", encoding="utf-8"))
        state = ckpt["generator"]
    model = HifiGanGenerator(config, state_dict=state, device=device)
    print(f"| Loaded model parameters from {file_path}.")
    print(f"| HifiGAN device: {device}.")
    return model, config, device


@register_vocoder
class HifiGAN(BaseVocoder):
    def __init__(self):
        base_dir = hparams["vocoder_ckpt"]
        config_path = f"{base_dir}/config.yaml"
        if os.path.exists(config_path):
            file_path = sorted(glob.glob(f"{base_dir}/model_ckpt_steps_*.*"), key=lambda x: int(
                re.findall(f"{base_dir}/model_ckpt_steps_(\\d+).*", x.replace("\\", "/"))[0]))[-1]
            print("| load HifiGAN: ", file_path)
            self.model, self.config, self.device = load_model(config_path=config_path, file_path=file_path)
        else:
            config_path = f"{base_dir}/config.json"
            file_path = f"{base_dir}/generator_v1"
            if os.path.exists(config_path):
                self.model, self.config, self.device = load_model(config_path=config_path, file_path=file_path)

    @classmethod
    def from_state_dict(cls, config, state_dict, device="cuda"):
        self = cls.__new__(cls)
        self.model, self.config, self.device = HifiGanGenerator(config, state_dict=state_dict, device=device), config, device
        return self

    def spec2wav(self, mel, **kwargs):
        with torch.no_grad():
            c = torch.FloatTensor(mel).unsqueeze(0).to(self.device)           # [1, T, 80], fed unscaled
            f0 = kwargs.get("f0")
            extra = {k: kwargs[k] for k in ("rand_ini", "sine_noise", "seed") if k in kwargs}
            if f0 is not None and hparams.get("use_nsf"):
                f0 = torch.FloatTensor(f0[None, :]).to(self.device)
                y = self.model.forward_mel(c, f0, 1.0, **extra).view(-1)
            else:
                y = self.model.forward_mel(c, None, 1.0).view(-1)
        wav_out = y.cpu().numpy()
        if hparams.get("vocoder_denoise_c", 0.0) > 0:
            try:   # host-side spectral denoise stays the reference's code (network/vocoders/vocoder_utils.py:7-15)
                from network.vocoders.vocoder_utils import denoise  # type: ignore
            except Exception as e:
                raise NotImplementedError("vocoder_denoise_c needs the reference's vocoder_utils (host code)") from e
            wav_out = denoise(wav_out, v=hparams["vocoder_denoise_c"])
        return wav_out

    @staticmethod
    def wav2spec(wav_fn, **kwargs):
        try:
            from network.vocoders.pwg import PWG  # type: ignore
        except Exception as e:
            raise NotImplementedError("wav2spec needs the reference's network.vocoders.pwg (host code)") from e
        return PWG.wav2spec(wav_fn, **kwargs)
