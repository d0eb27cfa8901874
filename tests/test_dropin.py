"""Drop-in wiring against the live reference tree (build container only): with the meta-path finder
installed, the reference's own `infer_tools.infer_tool` binds OUR classes without any edit."""
import os
import subprocess
import sys
import textwrap

import pytest

import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


def test_infer_tool_binds_native_classes(tmp_path):
    (tmp_path / "infer_tools").mkdir()
    (tmp_path / "infer_tools" / "f0_temp.json").write_text('{"info": "temp_dict"}')   # infer_tool.py:52 reads it relative to cwd
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import ref_harness as rh
        rh.install()                               # stubs for librosa etc. + set_hparams(config_nsf.yaml)
        import diffsvc_b200.dropin as dropin
        dropin.install()
        import infer_tools.infer_tool as it        # the reference's own, unmodified module
        import diffsvc_b200 as D
        from utils.hparams import hparams as ref_hparams
        assert it.GaussianDiffusion is D.GaussianDiffusion, it.GaussianDiffusion
        assert it.DiffNet is D.DiffNet
        assert D.hparams is ref_hparams            # one shared config dict (pndm_speedup side channel)
        from network.vocoders.base_vocoder import get_vocoder_cls, VOCODERS
        assert get_vocoder_cls(ref_hparams) is D.NsfHifiGAN, get_vocoder_cls(ref_hparams)
        assert VOCODERS["NsfHifiGAN"] is D.NsfHifiGAN
        import modules.nsf_hifigan.models as m
        assert m.load_model is D.vocoders.nsf_models.load_model
        import modules.fastspeech.fs2 as fs2
        dn = D.DiffNet(ref_hparams["audio_num_mel_bins"])
        gd = D.GaussianDiffusion(None, 128, dn, timesteps=ref_hparams["timesteps"], K_step=ref_hparams["K_step"],
                                 loss_type=ref_hparams["diff_loss_type"], spec_min=ref_hparams["spec_min"], spec_max=ref_hparams["spec_max"])
        assert isinstance(gd.fs2, fs2.FastSpeech2)  # conditioning stays the reference's own module
        import network.diff.diffusion as refd
        ref_keys = None
        dropin.uninstall()
        import network.diff.net as rnet, network.diff.diffusion as rdiff
        ref_gd = rdiff.GaussianDiffusion(None, 128, rnet.DiffNet(128), timesteps=1000, K_step=1000, loss_type="l2",
                                         spec_min=ref_hparams["spec_min"], spec_max=ref_hparams["spec_max"])
        a, b = gd.state_dict(), ref_gd.state_dict()
        assert set(a) == set(b), set(a) ^ set(b)   # strict load_ckpt compatibility
        assert all(tuple(a[k].shape) == tuple(b[k].shape) for k in a)
        gd.load_state_dict(b, strict=True)
        print("DROPIN_OK")
    """ % (ROOT, os.path.join(ROOT, "tests", "golden")))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_after_infer_patch_and_nvstft_alias(tmp_path):
    """install(patch_after_infer=True): the reference's own Svc gets the device-side after_infer the moment
    infer_tools.infer_tool is imported; modules.nsf_hifigan.nvSTFT resolves to the mel analysis kernel's host."""
    (tmp_path / "infer_tools").mkdir()
    (tmp_path / "infer_tools" / "f0_temp.json").write_text('{"info": "temp_dict"}')
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import ref_harness as rh
        rh.install()
        import diffsvc_b200.dropin as dropin
        dropin.install(patch_after_infer=True)
        import infer_tools.infer_tool as it
        from diffsvc_b200 import infer_glue
        assert it.Svc.after_infer is infer_glue.after_infer, it.Svc.after_infer
        assert it.Svc._dsvc_reference_after_infer.__module__ == "infer_tools.infer_tool"
        assert it.__file__.startswith(rh.REFERENCE_ROOT), it.__file__      # still the reference's own module
        import modules.nsf_hifigan.nvSTFT as nv
        import diffsvc_b200.vocoders.nvstft as ours
        assert nv is ours and nv.STFT is ours.STFT
        # late patching of an already imported module
        dropin.uninstall()
        sys.modules.pop("infer_tools.infer_tool")
        import infer_tools.infer_tool as it2
        assert it2.Svc.after_infer is not infer_glue.after_infer
        dropin.install(patch_after_infer=True)
        assert it2.Svc.after_infer is infer_glue.after_infer
        print("PATCH_OK")
    """ % (ROOT, os.path.join(ROOT, "tests", "golden")))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert "PATCH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_pitch_extractor_alias_and_keys(tmp_path):
    """The reference's infer_tool binds our PitchExtractor, whose state_dict equals the reference module's."""
    (tmp_path / "infer_tools").mkdir()
    (tmp_path / "infer_tools" / "f0_temp.json").write_text('{"info": "temp_dict"}')
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import ref_harness as rh
        rh.install()
        import diffsvc_b200.dropin as dropin
        dropin.install()
        import infer_tools.infer_tool as it
        import diffsvc_b200 as D
        assert it.PitchExtractor is D.PitchExtractor
        ours = D.PitchExtractor().state_dict()
        dropin.uninstall()
        import modules.fastspeech.pe as ref_pe
        assert ref_pe.PitchExtractor is not D.PitchExtractor
        ref = ref_pe.PitchExtractor().state_dict()
        assert set(ours) == set(ref), set(ours) ^ set(ref)
        assert all(tuple(ours[k].shape) == tuple(ref[k].shape) for k in ref)
        print("PE_OK")
    """ % (ROOT, os.path.join(ROOT, "tests", "golden")))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert "PE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
