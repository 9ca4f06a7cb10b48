from .._impl import randn_tensor  # noqa: F401
