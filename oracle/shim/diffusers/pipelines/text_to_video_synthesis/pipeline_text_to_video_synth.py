from ..._impl import TextToVideoSDPipeline, TextToVideoSDPipelineOutput, tensor2vid  # noqa: F401
