from .._impl import Downsample2D, ResnetBlock2D, TemporalConvLayer, Upsample2D  # noqa: F401
