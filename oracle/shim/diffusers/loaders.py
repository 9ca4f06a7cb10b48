from ._impl import LoraLoaderMixin, TextualInversionLoaderMixin  # noqa: F401
