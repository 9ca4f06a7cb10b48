"""ORACLE shim: a stand-in for `diffusers==0.24.0` (see _impl.py header; parity unpinned). Test infrastructure only."""
from ._impl import (AutoencoderKL, DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler,  # noqa: F401
                    DiffusionPipeline, StableVideoDiffusionPipeline, TextToVideoSDPipeline)
__version__ = "0.24.0+oracle-shim"
