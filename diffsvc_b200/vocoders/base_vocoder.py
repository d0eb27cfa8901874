"""Vocoder registry / plug-in API: network/vocoders/base_vocoder.py:2-39."""
import importlib

try:  # drop-in: share the reference's registry so get_vocoder_cls() in infer_tool finds our classes
    from network.vocoders.base_vocoder import VOCODERS  # type: ignore
except Exception:
    VOCODERS = {}


def register_vocoder(cls):
    VOCODERS[cls.__name__.lower()] = cls
    VOCODERS[cls.__name__] = cls
    return cls


def get_vocoder_cls(hparams):
    if hparams["vocoder"] in VOCODERS:
        return VOCODERS[hparams["vocoder"]]
    vocoder_cls = hparams["vocoder"]
    pkg = ".".join(vocoder_cls.split(".")[:-1])
    cls_name = vocoder_cls.split(".")[-1]
    return getattr(importlib.import_module(pkg), cls_name)


class BaseVocoder:
    def spec2wav(self, mel):
        """:param mel: [T, 80]  :return: wav: [T']"""
        raise NotImplementedError

    @staticmethod
    def wav2spec(wav_fn):
        """:param wav_fn: str  :return: wav, mel: [T, 80]"""
        raise NotImplementedError
