from .._impl import TransformerTemporalModel, TransformerTemporalModelOutput  # noqa: F401
