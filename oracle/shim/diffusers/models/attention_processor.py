from .._impl import Attention, AttnProcessor2_0  # noqa: F401
