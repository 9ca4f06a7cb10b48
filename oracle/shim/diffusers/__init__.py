"""ORACLE shim: a stand-in for `diffusers==0.24.0` (see _impl.py header; parity unpinned). Test infrastructure only."""
from ._impl import (AutoencoderKL, DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler,  # noqa: F401
                    DiffusionPipeline, StableVideoDiffusionPipeline, TextToVideoSDPipeline,
                    AutoencoderKLTemporalDecoder, EulerDiscreteScheduler, UNetSpatioTemporalConditionModel)
import PIL.Image  # noqa: E402,F401  (real diffusers imports it; models/pipeline.py:334 relies on `PIL.Image` being loaded)
__version__ = "0.24.0+oracle-shim"
