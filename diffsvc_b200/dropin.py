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
"""
import importlib
import importlib.abc
import importlib.util
import runpy
import sys

ALIASES = {
    "network.diff.net": "diffsvc_b200.net",
    "network.diff.diffusion": "diffsvc_b200.diffusion",
    "network.vocoders.nsf_hifigan": "diffsvc_b200.vocoders.nsf_hifigan",
    "modules.nsf_hifigan.models": "diffsvc_b200.vocoders.nsf_models",
    "network.vocoders.hifigan": "diffsvc_b200.vocoders.hifigan",
}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)     # the very same module object under a second name

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        tgt = ALIASES.get(fullname)
        if tgt is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(tgt))


_installed = None


def install():
    global _installed
    if _installed is None:
        _installed = _Finder()
        sys.meta_path.insert(0, _installed)
        for name in ALIASES:               # drop stale reference copies imported before install()
            sys.modules.pop(name, None)
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
