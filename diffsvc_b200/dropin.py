"""Drop-in wiring: serve the B200-native implementations under the reference's own module names, so
`inference.ipynb`, `infer.py`, `batch.py` and `flask_api.py` run UNCHANGED.

The reference's scripts import by absolute module name from the repo root
(`from network.diff.diffusion import GaussianDiffusion`, infer_tools/infer_tool.py:16-18), so a
`sys.meta_path` finder placed first wins regardless of `sys.path` order and leaves every other
reference module (HuBERT, f0, slicer, fs2, hparams ...) untouched:

    import diffsvc_b200.dropin; diffsvc_b200.dropin.install()      # e.g. from sitecustomize / a launcher
    # or:  python -m diffsvc_b200.dropin infer.py ...

Replaced modules -> ours:
    network.diff.net               -> diffsvc_b200.net        (DiffNet)
    network.diff.diffusion         -> diffsvc_b200.diffusion  (GaussianDiffusion, OfflineGaussianDiffusion)
    network.vocoders.nsf_hifigan   -> diffsvc_b200.vocoders.nsf_hifigan (NsfHifiGAN, registered in the reference's VOCODERS)
    modules.nsf_hifigan.models     -> diffsvc_b200.vocoders.nsf_models  (load_model, Generator)
    network.vocoders.hifigan       -> diffsvc_b200.vocoders.hifigan     (HifiGAN, load_model; 24 kHz models)
    modules.nsf_hifigan.nvSTFT     -> diffsvc_b200.vocoders.nvstft      (STFT, load_wav_to_torch; mel analysis kernel)
    modules.fastspeech.pe          -> diffsvc_b200.pe         (PitchExtractor; inference only -- training/pe.py needs the reference's)

`install(patch_after_infer=True)` (or DSVC_PATCH_AFTER_INFER=1) additionally replaces `Svc.after_infer`
(infer_tools/infer_tool.py:172-200) by diffsvc_b200.infer_glue.after_infer the moment infer_tools.infer_tool is
imported, which keeps the denoised mel on the device between the sampler and the vocoder.
"""
import importlib
import importlib.abc
import importlib.util
import os
import runpy
import sys

ALIASES = {
    "network.diff.net": "diffsvc_b200.net",
    "network.diff.diffusion": "diffsvc_b200.diffusion",
    "network.vocoders.nsf_hifigan": "diffsvc_b200.vocoders.nsf_hifigan",
    "modules.nsf_hifigan.models": "diffsvc_b200.vocoders.nsf_models",
    "network.vocoders.hifigan": "diffsvc_b200.vocoders.hifigan",
    "modules.nsf_hifigan.nvSTFT": "diffsvc_b200.vocoders.nvstft",
    "modules.fastspeech.pe": "diffsvc_b200.pe",
}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)     # the very same module object under a second name

    def exec_module(self, module):
        pass


class _PatchingLoader(importlib.abc.Loader):
    """Runs the reference's own loader for infer_tools.infer_tool, then rebinds Svc.after_infer."""

    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        from . import infer_glue
        if hasattr(module, "Svc"):
            infer_glue.patch(module.Svc)


class _Finder(importlib.abc.MetaPathFinder):
    patch_after_infer = False

    def find_spec(self, fullname, path, target=None):
        tgt = ALIASES.get(fullname)
        if tgt is not None:
            return importlib.util.spec_from_loader(fullname, _AliasLoader(tgt))
        if self.patch_after_infer and fullname == "infer_tools.infer_tool":
            for finder in sys.meta_path:         # the finder that would have served it without us
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(fullname, path, target)
                if spec is not None and spec.loader is not None:
                    spec.loader = _PatchingLoader(spec.loader)
                    return spec
        return None


_installed = None


def install(patch_after_infer=None):
    global _installed
    if patch_after_infer is None:
        patch_after_infer = os.environ.get("DSVC_PATCH_AFTER_INFER") == "1"
    if _installed is None:
        _installed = _Finder()
        sys.meta_path.insert(0, _installed)
        for name in ALIASES:               # drop stale reference copies imported before install()
            sys.modules.pop(name, None)
    _installed.patch_after_infer = bool(patch_after_infer)
    if patch_after_infer and "infer_tools.infer_tool" in sys.modules:    # already imported: patch in place
        from . import infer_glue
        infer_glue.patch(sys.modules["infer_tools.infer_tool"].Svc)
    return _installed


def uninstall():
    global _installed
    if _installed is not None:
        sys.meta_path.remove(_installed)
        for name in ALIASES:
            sys.modules.pop(name, None)
        _installed = None


if __name__ == "__main__":                 # python -m diffsvc_b200.dropin infer.py [args]
    install()
    sys.argv = sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
