"""Overrides the reference module of the same name (host-side UniPC loop, bit-exact vs the reference)."""
from videocof_b200.scheduler import FlowUniPCMultistepScheduler, SchedulerOutput  # noqa: F401
