from .._impl import AutoencoderKL  # noqa: F401
