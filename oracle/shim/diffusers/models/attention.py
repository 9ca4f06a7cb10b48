from .._impl import BasicTransformerBlock, FeedForward, GEGLU  # noqa: F401
