from . import hifigan  # noqa: F401  (registers HifiGAN and NsfHifiGAN, as network/vocoders/__init__.py does)
from . import nsf_hifigan  # noqa: F401
