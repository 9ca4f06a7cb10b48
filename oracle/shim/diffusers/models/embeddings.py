from .._impl import TimestepEmbedding, Timesteps, get_timestep_embedding  # noqa: F401
