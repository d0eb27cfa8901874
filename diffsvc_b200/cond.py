"""Conditioning encoder used when the reference's `modules.fastspeech.fs2.FastSpeech2` is not
importable (stand-alone tests / bench): the `no_fs2` path of FastSpeech2.forward
(modules/fastspeech/fs2.py:94-154, add_pitch :185-238, utils/pitch_utils.py:17-31,63-76).
SURVEY.md section 8(f) row 1: CUDA tensors go through one fused kernel (`dsvc_cond_encode`); the
PyTorch restatement below is the host-side mirror for CPU tensors (state-dict plumbing and CPU tests).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .hparams import hparams


def f0_to_coarse(f0, hp):
    f0_bin, f0_max, f0_min = hp["f0_bin"], hp["f0_max"], hp["f0_min"]
    mel_min = 1127 * np.log(1 + f0_min / 700)
    mel_max = 1127 * np.log(1 + f0_max / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - mel_min) * (f0_bin - 2) / (mel_max - mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > f0_bin - 1] = f0_bin - 1
    return (f0_mel + 0.5).long()


class CondEncoder(nn.Module):
    """State-dict compatible with the reference for the one tensor this path reads: `pitch_embed.weight`."""

    def __init__(self, dictionary=None, out_dims=None):
        super().__init__()
        self.hidden_size = hparams["hidden_size"]
        self.padding_idx = 0
        self.pitch_embed = nn.Embedding(300, self.hidden_size, self.padding_idx)
        nn.init.normal_(self.pitch_embed.weight, mean=0, std=self.hidden_size ** -0.5)
        nn.init.constant_(self.pitch_embed.weight[self.padding_idx], 0)

    def forward(self, hubert, mel2ph=None, spk_embed=None, ref_mels=None, f0=None, uv=None, energy=None,
                skip_decoder=True, spk_embed_dur_id=None, spk_embed_f0_id=None, infer=False, **kwargs):
        # a missing `no_fs2` means "encoder enabled" in the reference (fs2.py:98): not this stand-alone module's job
        if not hparams.get("no_fs2", False) or hparams.get("use_spk_embed") or hparams.get("use_spk_id") \
                or hparams.get("use_energy_embed") or hparams.get("pitch_norm", "log") != "log":
            raise NotImplementedError("stand-alone CondEncoder covers the config_nsf.yaml conditioning only; "
                                      "run inside the reference tree to use its FastSpeech2")
        ret = {"mel2ph": mel2ph}
        if hubert.is_cuda:      # device path: one fused kernel through the C-ABI (include/dsvc.h: dsvc_cond_encode)
            from . import _lib
            hub = hubert.detach().to(torch.float32).contiguous()
            m2p = mel2ph.to(torch.int64).contiguous()
            f0c = f0.detach().to(torch.float32).contiguous()
            emb = self.pitch_embed.weight.detach().to(hub.device, torch.float32).contiguous()
            B, Th, H = hub.shape
            T = m2p.shape[1]
            out = torch.empty(B, T, H, device=hub.device, dtype=torch.float32)
            f0d = torch.empty(B, T, device=hub.device, dtype=torch.float32)
            with torch.cuda.device(hub.device):
                _lib.check(_lib.load().dsvc_cond_encode(_lib.dptr(hub), _lib.dptr(m2p), _lib.dptr(f0c), _lib.dptr(emb), B, Th, T, H,
                                                        int(hparams["f0_bin"]), float(hparams["f0_min"]), float(hparams["f0_max"]),
                                                        _lib.dptr(out), _lib.dptr(f0d), _lib.current_stream()))
            f0[mel2ph == 0] = 0                                # fs2.py:226-227 (in-place on the caller's tensor)
            pitch_pad = None
            ret["f0_denorm"] = f0d
            ret["decoder_inp"] = out
            return ret
        decoder_inp = F.pad(hubert, [0, 0, 1, 0])
        mel2ph_ = mel2ph[..., None].repeat([1, 1, hubert.shape[-1]])
        decoder_inp = torch.gather(decoder_inp, 1, mel2ph_)
        tgt_nonpadding = (mel2ph > 0).float()[:, :, None]
        pitch_padding = (mel2ph == 0)
        f0_denorm = 2 ** f0
        f0_denorm[pitch_padding] = 0
        ret["f0_denorm"] = f0_denorm
        f0[pitch_padding] = 0                                  # fs2.py:226-227 (in-place on the caller's tensor)
        pitch = f0_to_coarse(f0_denorm, hparams)
        ret["pitch_pred"] = pitch.unsqueeze(-1)
        ret["decoder_inp"] = (decoder_inp + self.pitch_embed(pitch)) * tgt_nonpadding
        return ret
