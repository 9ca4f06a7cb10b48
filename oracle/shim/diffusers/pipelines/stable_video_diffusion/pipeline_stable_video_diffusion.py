from ..._impl import StableVideoDiffusionPipeline, StableVideoDiffusionPipelineOutput, tensor2vid  # noqa: F401
