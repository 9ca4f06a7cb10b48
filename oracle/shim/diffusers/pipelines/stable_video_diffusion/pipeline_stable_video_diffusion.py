from ..._impl import StableVideoDiffusionPipeline, StableVideoDiffusionPipelineOutput  # noqa: F401
from ..._impl import svd_tensor2vid as tensor2vid  # noqa: F401
