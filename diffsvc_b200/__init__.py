"""diffsvc_b200 -- B200-native (sm_100a) inference hot path of diffusion-SVC behind the reference's
own Python API: `DiffNet`, `GaussianDiffusion`, `network.vocoders` registry / `NsfHifiGAN`.

Host code is Python over PyTorch tensors (device memory, streams); all compute on the hot path is
hand-written CUDA in csrc/, reached through the C-ABI of include/dsvc.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .hparams import hparams, set_hparams  # noqa: F401
from .net import DiffNet  # noqa: F401
from .diffusion import GaussianDiffusion, OfflineGaussianDiffusion  # noqa: F401
from .vocoders.base_vocoder import VOCODERS, BaseVocoder, get_vocoder_cls, register_vocoder  # noqa: F401
from .vocoders.nsf_hifigan import NsfHifiGAN  # noqa: F401
from .vocoders.hifigan import HifiGAN, HifiGanGenerator  # noqa: F401
from .vocoders.nvstft import STFT  # noqa: F401
from .pe import PitchExtractor  # noqa: F401
from . import infer_glue  # noqa: F401

__version__ = "0.1.0"
