"""`PitchExtractor` -- drop-in for modules/fastspeech/pe.py:120-149 backed by libdsvc (`dsvc_pe_forward`).

Same constructor (`PitchExtractor(n_mel_bins=80, conv_layers=2)`, sizes from `hparams`), same parameter / buffer
names and shapes as the reference module tree, so `utils.load_ckpt(self.pe, hparams['pe_ckpt'], 'model',
strict=True)` and `.cuda()` (infer_tools/infer_tool.py:134-136) work unchanged; `forward(mel_input [B, T, 80])`
returns the same dict (`pitch_pred [B, T, 2]`, `f0_denorm_pred [B, T]`).  The torch sub-modules below only HOLD the
fp32 parameters -- no torch op runs in `forward`, and there is no CPU path.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from .hparams import hparams


def _sinusoid_table(rows, dim):
    """SinusoidalPositionalEmbedding.get_embedding(rows, dim, padding_idx=0), modules/commons/common_layers.py:105-122,
    with the reference's own float ops on the host (it is a constant of the module, not per-call work)."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
    ang = torch.arange(rows, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
    table = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(rows, -1)
    if dim % 2 == 1:
        table = torch.cat([table, torch.zeros(rows, 1)], dim=1)
    table[0, :] = 0
    return table


class _Positions(nn.Module):      # key `pitch_predictor.embed_positions._float_tensor` (common_layers.py:103)
    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.FloatTensor(1))


class _Holder(nn.Module):
    """A bare namespace module: children are attached by the builder below."""


def _prenet(in_dim, out_dim, kernel, n_layers):          # pe.py:8-21
    m = _Holder()
    layers, d = [], in_dim
    for _ in range(n_layers):
        layers.append(nn.Sequential(nn.Conv1d(d, out_dim, kernel_size=kernel, padding=kernel // 2), nn.ReLU(),
                                    nn.BatchNorm1d(out_dim)))
        d = out_dim
    m.layers = nn.ModuleList(layers)
    m.out_proj = nn.Linear(out_dim, out_dim)
    return m


def _xavier_linear(i, o):                                # common_layers.py:80-85
    m = nn.Linear(i, o)
    nn.init.xavier_uniform_(m.weight)
    nn.init.constant_(m.bias, 0.)
    return m


def _conv_stacks(n_chans, n_layers, kernel):             # pe.py:83-98 with norm='gn'
    m = _Holder()
    m.in_proj = _xavier_linear(n_chans, n_chans)
    blocks = []
    for _ in range(n_layers):
        blk = _Holder()
        blk.conv = _Holder()                             # ConvNorm wraps its Conv1d as `.conv` (common_layers.py:49)
        blk.conv.conv = nn.Conv1d(n_chans, n_chans, kernel, padding=kernel // 2)
        nn.init.xavier_uniform_(blk.conv.conv.weight)
        blk.norm = nn.GroupNorm(n_chans // 16, n_chans)
        blocks.append(blk)
    m.conv = nn.ModuleList(blocks)
    m.out_proj = _xavier_linear(n_chans, n_chans)
    return m


def _pitch_predictor(idim, n_chans, n_layers, kernel, odim):     # tts_modules.py:192-220
    m = _Holder()
    convs = []
    for i in range(n_layers):
        convs.append(nn.Sequential(nn.Identity(), nn.Conv1d(idim if i == 0 else n_chans, n_chans, kernel), nn.ReLU(),
                                   nn.LayerNorm(n_chans, eps=1e-12), nn.Identity()))
    m.conv = nn.ModuleList(convs)
    m.linear = nn.Linear(n_chans, odim)
    m.embed_positions = _Positions()
    m.pos_embed_alpha = nn.Parameter(torch.Tensor([1]))
    return m


class PitchExtractor(nn.Module):
    def __init__(self, n_mel_bins=80, conv_layers=2):
        super().__init__()
        self.n_mel_bins = n_mel_bins
        self.hidden_size = hparams["hidden_size"]
        self.predictor_hidden = hparams["predictor_hidden"] if hparams.get("predictor_hidden", -1) > 0 else self.hidden_size
        self.conv_layers = conv_layers
        self.predictor_kernel = hparams.get("predictor_kernel", 5)
        self.pad_same = hparams.get("ffn_padding", "SAME") == "SAME"
        self.mel_prenet = _prenet(n_mel_bins, self.hidden_size, 5, 3)
        if conv_layers > 0:
            self.mel_encoder = _conv_stacks(self.hidden_size, conv_layers, 5)
        self.pitch_predictor = _pitch_predictor(self.hidden_size, self.predictor_hidden, 5, self.predictor_kernel, 2)
        self._h = None
        self._h_key = None
        self._pos_rows = 4096                                    # init_size (tts_modules.py:219), grown on demand

    # ---- native handle -----------------------------------------------------------------------------------------
    def _key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())) + (
            self._pos_rows, hparams.get("pitch_norm"), hparams.get("pitch_type"), hparams.get("use_uv"),
            hparams.get("f0_mean"), hparams.get("f0_std"))

    def release(self):
        if self._h is not None:
            _lib.load().dsvc_pe_destroy(self._h)
        self._h = self._h_key = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def handle(self):
        key = self._key()
        if self._h is not None and key == self._h_key:
            return self._h
        self.release()
        lib = _lib.load()
        if not torch.cuda.is_available() or lib.dsvc_device_count() == 0:
            raise _lib.DsvcError("diffsvc_b200.PitchExtractor needs an sm_100 (B200) device: there is no CPU fallback")
        keep = []

        def f(t):
            t = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            return _lib.fptr(t)

        def fa(ts):
            ts = [t.detach().to("cpu", torch.float32).contiguous() for t in ts]
            arr = _lib.fptr_array(ts)
            keep.extend(ts + [arr])
            return arr

        pre, pp = self.mel_prenet, self.pitch_predictor
        w = _lib.PeWeights()
        w.prenet_conv_w, w.prenet_conv_b = fa([l[0].weight for l in pre.layers]), fa([l[0].bias for l in pre.layers])
        w.prenet_bn_w, w.prenet_bn_b = fa([l[2].weight for l in pre.layers]), fa([l[2].bias for l in pre.layers])
        w.prenet_bn_mean = fa([l[2].running_mean for l in pre.layers])
        w.prenet_bn_var = fa([l[2].running_var for l in pre.layers])
        w.prenet_out_w, w.prenet_out_b = f(pre.out_proj.weight), f(pre.out_proj.bias)
        if self.conv_layers > 0:
            enc = self.mel_encoder
            w.enc_in_w, w.enc_in_b = f(enc.in_proj.weight), f(enc.in_proj.bias)
            w.enc_conv_w, w.enc_conv_b = fa([b.conv.conv.weight for b in enc.conv]), fa([b.conv.conv.bias for b in enc.conv])
            w.enc_gn_w, w.enc_gn_b = fa([b.norm.weight for b in enc.conv]), fa([b.norm.bias for b in enc.conv])
            w.enc_out_w, w.enc_out_b = f(enc.out_proj.weight), f(enc.out_proj.bias)
        w.pred_conv_w, w.pred_conv_b = fa([s[1].weight for s in pp.conv]), fa([s[1].bias for s in pp.conv])
        w.pred_ln_w, w.pred_ln_b = fa([s[3].weight for s in pp.conv]), fa([s[3].bias for s in pp.conv])
        w.pred_linear_w, w.pred_linear_b = f(pp.linear.weight), f(pp.linear.bias)
        w.pos_table = f(_sinusoid_table(self._pos_rows, self.hidden_size))
        w.pos_embed_alpha = f(pp.pos_embed_alpha)
        norm = {"log": 1, "standard": 2}.get(hparams.get("pitch_norm", "log"), 0)
        apply_uv = int(hparams.get("pitch_type", "frame") == "frame" and bool(hparams.get("use_uv", False)))
        cfg = _lib.PeConfig(self.n_mel_bins, self.hidden_size, self.predictor_hidden, 3, 5, self.conv_layers, 5,
                            self.hidden_size // 16, 5, self.predictor_kernel, int(self.pad_same), 2, self._pos_rows, norm,
                            apply_uv, float(hparams.get("f0_mean", 0.0) or 0.0), float(hparams.get("f0_std", 1.0) or 1.0),
                            1e-5, 1e-5, 1e-12)
        h = C.c_void_p()
        dev = pp.linear.weight.device
        with torch.cuda.device(dev if dev.type == "cuda" else torch.cuda.current_device()):
            _lib.check(lib.dsvc_pe_create(C.byref(h), C.byref(cfg), C.byref(w), _lib.current_stream()))
        self._h, self._h_key = h, key
        return h

    def forward(self, mel_input=None):
        """mel_input: CUDA fp32 [B, T, n_mel_bins] -> {'pitch_pred': [B, T, 2], 'f0_denorm_pred': [B, T]}"""
        if not mel_input.is_cuda:
            raise _lib.DsvcError("PitchExtractor.forward: CUDA tensors only (no CPU path)")
        mel = mel_input.detach().to(torch.float32).contiguous()
        B, T, M = mel.shape
        assert M == self.n_mel_bins, (mel.shape, self.n_mel_bins)
        if B == 0 or T == 0:
            return {"pitch_pred": mel.new_empty(B, T, 2), "f0_denorm_pred": mel.new_empty(B, T)}
        if T + 1 > self._pos_rows:               # common_layers.py:127-134 grows the table the same way
            self._pos_rows = T + 1
        h = self.handle()
        pred = torch.empty(B, T, 2, device=mel.device, dtype=torch.float32)
        f0 = torch.empty(B, T, device=mel.device, dtype=torch.float32)
        with torch.cuda.device(mel.device):
            _lib.check(_lib.load().dsvc_pe_forward(h, _lib.dptr(mel), B, T, _lib.dptr(pred), _lib.dptr(f0), _lib.current_stream()))
        return {"pitch_pred": pred, "f0_denorm_pred": f0}
