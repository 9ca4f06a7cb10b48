from .._impl import BaseOutput, logging  # noqa: F401
