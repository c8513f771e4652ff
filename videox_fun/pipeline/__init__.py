"""videox_fun.pipeline — WanPipeline with the reference's __call__ contract (pipeline_wan.py:518-799)."""
from videocof_b200.pipeline import WanPipeline, WanPipelineOutput

WanFunPipeline = WanPipeline
