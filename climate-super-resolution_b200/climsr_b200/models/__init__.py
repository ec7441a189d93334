from .esrgan import ESRGANGenerator  # noqa: F401
