"""`Svc.after_infer` with the tensors kept on the device.

The reference's `Svc.after_infer` (infer_tools/infer_tool.py:172-200) copies every tensor of the prediction to
the host, masks padding frames and clips the mel in numpy, then hands numpy arrays to `vocoder.spec2wav`, which
copies them back to the GPU (network/vocoders/nsf_hifigan.py:62-72).  Here the denoised mel and the predicted f0
stay in HBM: one `dsvc_compact_frames` launch does the mask + clip + f0 selection, the vocoder consumes its
output directly, and only the results the caller receives (f0_gt, f0_pred, waveform) are copied to the host.

`after_infer(self, prediction, singer, in_path)` has the reference's signature and return value, so it can be
bound in place of `Svc.after_infer` (`diffsvc_b200.dropin.install(patch_after_infer=True)`).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .hparams import hparams   # the reference's own dict when its tree is importable (hparams.py)


def compact_frames(mel, f0, vmin, vmax):
    """mel [T, M], f0 [T] or None (CUDA fp32) -> (mel_kept [N, M] clipped to [vmin, vmax], f0_kept [N] or None).

    Kept frames are those with abs(mel).sum(-1) > 0 (infer_tool.py:183-186,193).  One kernel + a 4-byte D2H (N)."""
    if not mel.is_cuda:
        raise _lib.DsvcError("compact_frames: CUDA tensors only (no CPU path)")
    lib = _lib.load()
    mel = mel.contiguous().float()
    T, M = mel.shape
    mel_out = torch.empty_like(mel)
    n_kept = torch.zeros(1, dtype=torch.int32, device=mel.device)
    if f0 is not None:
        f0 = f0.contiguous().float()
        assert f0.shape == (T,), (f0.shape, T)
        f0_out = torch.empty_like(f0)
    if T > 0:
        with torch.cuda.device(mel.device):
            _lib.check(lib.dsvc_compact_frames(_lib.dptr(mel), _lib.dptr(f0) if f0 is not None else None, T, M,
                                               C.c_float(vmin), C.c_float(vmax), _lib.dptr(mel_out),
                                               _lib.dptr(f0_out) if f0 is not None else None, _lib.dptr(n_kept),
                                               _lib.current_stream()))
    n = int(n_kept.item())
    return mel_out[:n], (f0_out[:n] if f0 is not None else None)


def _np(v):
    return v.cpu().numpy() if isinstance(v, torch.Tensor) else v


def after_infer(self, prediction, singer, in_path):
    """Drop-in for Svc.after_infer (infer_tool.py:172-200); B = 1 like the reference's boolean indexing."""
    mel_pred = prediction["outputs"]
    f0_pred = prediction.get("f0_pred")
    on_device = isinstance(mel_pred, torch.Tensor) and mel_pred.is_cuda and hasattr(self.vocoder, "spec2wav_device") \
        and isinstance(f0_pred, torch.Tensor)
    if not on_device:                            # not our tensors / vocoder: the reference's own host path
        return type(self)._dsvc_reference_after_infer(self, prediction, singer, in_path)

    mel_gt = _np(prediction["mels"])
    mel_gt_mask = np.abs(mel_gt).sum(-1) > 0
    f0_gt = _np(prediction.get("f0_gt"))
    f0_gt = f0_gt[mel_gt_mask]                   # f0_pred is not None here

    mel_dev = mel_pred.reshape(-1, mel_pred.shape[-1])
    f0_dev = f0_pred.to(mel_dev.device).reshape(-1)[: mel_dev.shape[0]]
    if f0_dev.shape[0] < mel_dev.shape[0]:
        raise ValueError("f0_pred has fewer frames (%d) than the mel (%d)" % (f0_dev.shape[0], mel_dev.shape[0]))
    mel_kept, f0_kept = compact_frames(mel_dev, f0_dev, float(hparams["mel_vmin"]), float(hparams["mel_vmax"]))

    for k, v in list(prediction.items()):        # callers read the dict afterwards: same host-side types as the reference
        if type(v) is torch.Tensor:
            prediction[k] = v.cpu().numpy()
    f0_host = f0_kept.cpu().numpy()
    if singer:
        data_path = in_path.replace("batch", "singer_data")
        np.save(data_path[:-4] + "_mel.npy", mel_kept.cpu().numpy())
        np.save(data_path[:-4] + "_f0.npy", f0_host)
    wav_pred = self.vocoder.spec2wav_device(mel_kept, f0_kept).cpu().numpy()
    return f0_gt, f0_host, wav_pred


def patch(svc_cls):
    """Bind `after_infer` over `svc_cls.after_infer`, keeping the original for inputs that are not on the device."""
    if getattr(svc_cls, "_dsvc_reference_after_infer", None) is None:
        svc_cls._dsvc_reference_after_infer = svc_cls.after_infer
        svc_cls.after_infer = after_infer
    return svc_cls
