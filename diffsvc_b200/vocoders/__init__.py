from . import nsf_hifigan  # noqa: F401  (registers NsfHifiGAN, as network/vocoders/__init__.py does)
