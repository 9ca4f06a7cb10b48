from .._impl import Transformer2DModel, Transformer2DModelOutput  # noqa: F401
