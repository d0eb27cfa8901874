"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/dsvc.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from diffsvc_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "dsvc.h")).read()
    declared = set(re.findall(r"\b(dsvc_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dsvc_version()
    assert lib.dsvc_abi() == _lib.header_crc()          # the library was built from THIS include/dsvc.h


def test_struct_layouts_match_header():
    from diffsvc_b200 import _lib
    assert C.sizeof(_lib.DiffnetConfig) == 7 * 4
    assert C.sizeof(_lib.DiffnetWeights) == 19 * 8
    assert C.sizeof(_lib.NsfConfig) == 4 * (4 + 8 + 8 + 1 + 8 + 1 + 64 + 1 + 1)
    assert C.sizeof(_lib.NsfWeights) == 14 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    import diffsvc_b200 as D
    from diffsvc_b200 import _lib
    lib = _lib.load()
    assert lib.dsvc_device_count() == 0
    h = C.c_void_p()
    cfg = _lib.DiffnetConfig(128, 384, 256, 20, 4, 1000, 0)
    w = _lib.DiffnetWeights()
    rc = lib.dsvc_diffnet_create(C.byref(h), C.byref(cfg), C.byref(w), None)
    assert rc == -3 and b"no CPU fallback" in lib.dsvc_last_error()
    with pytest.raises(_lib.DsvcError):
        D.DiffNet(128).handle()
    # every other compute entry point refuses as well (dummy non-null pointers: the device probe comes first)
    buf = (C.c_float * 4096)()
    ptr = C.cast(buf, C.c_void_p)
    mcfg = _lib.MelConfig(64, 16, 8, 1e-5, 1.0)
    assert lib.dsvc_mel_analysis(C.byref(mcfg), ptr, 1024, ptr, ptr, None, None, ptr, None) == -3
    assert lib.dsvc_compact_frames(ptr, None, 4, 8, -6.0, 1.5, ptr, None, ptr, None) == -3
    assert lib.dsvc_cond_encode(ptr, ptr, ptr, ptr, 1, 4, 4, 8, 256, 40.0, 1100.0, ptr, ptr, None) == -3
    ncfg, nw = _lib.NsfConfig(), _lib.NsfWeights()
    ncfg.num_upsamples, ncfg.num_kernels, ncfg.num_dilations, ncfg.harmonic_num = 1, 1, 1, 8
    assert lib.dsvc_nsf_create(C.byref(h), C.byref(ncfg), C.byref(nw), None) == -3
    pcfg, pw = _lib.PeConfig(), _lib.PeWeights()
    assert lib.dsvc_pe_create(C.byref(h), C.byref(pcfg), C.byref(pw), None) == -3
    assert b"no CPU fallback" in lib.dsvc_last_error()
    with pytest.raises(_lib.DsvcError):
        D.PitchExtractor().handle()


def test_state_dict_keys_match_reference_layout():
    """Same parameter names / shapes as network/diff/net.py + the 12 schedule buffers of diffusion.py."""
    import diffsvc_b200 as D
    from oracle import diffsvc_oracle as O
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    hparams.clear(); hparams.update(DEFAULTS_44K)
    dn = D.DiffNet(128)
    ref = O.synth_diffnet_weights()
    sd = dn.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0])
    keys = set(gd.state_dict())
    assert set(O.SCHEDULE_KEYS) <= keys and {"spec_min", "spec_max", "fs2.pitch_embed.weight"} <= keys
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(getattr(gd, k), sched[k]), k


def test_vocoder_registry():
    import diffsvc_b200 as D
    assert D.VOCODERS["NsfHifiGAN"] is D.NsfHifiGAN and D.VOCODERS["nsfhifigan"] is D.NsfHifiGAN
    assert D.get_vocoder_cls({"vocoder": "NsfHifiGAN"}) is D.NsfHifiGAN
    assert D.get_vocoder_cls({"vocoder": "diffsvc_b200.vocoders.nsf_hifigan.NsfHifiGAN"}) is D.NsfHifiGAN


def test_cond_encoder_matches_oracle():
    import diffsvc_b200 as D
    from diffsvc_b200.cond import CondEncoder
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    from oracle import diffsvc_oracle as O
    hparams.clear(); hparams.update(DEFAULTS_44K)
    enc = CondEncoder()
    g = torch.Generator().manual_seed(0)
    hub = torch.randn(2, 30, 256, generator=g)
    mel2ph = torch.randint(1, 31, (2, 50), generator=g).sort(dim=1).values
    mel2ph[1, 45:] = 0
    f0 = torch.log2(torch.rand(2, 50, generator=g) * 600 + 50)
    ret = enc(hub, mel2ph, None, None, f0.clone(), None, None)
    dec, f0d = O.cond_encoder(enc.pitch_embed.weight.detach(), hub, mel2ph, f0.clone())
    assert torch.equal(ret["decoder_inp"], dec) and torch.equal(ret["f0_denorm"], f0d)
