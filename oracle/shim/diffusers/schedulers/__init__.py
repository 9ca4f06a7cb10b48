from .._impl import DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler  # noqa: F401
