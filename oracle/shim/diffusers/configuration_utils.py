from ._impl import ConfigMixin, FrozenDict, register_to_config  # noqa: F401
