from videocof_b200.pipeline import WanPipeline, WanPipelineOutput, randn_tensor  # noqa: F401
