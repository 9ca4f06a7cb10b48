from .._impl import ModelMixin  # noqa: F401
