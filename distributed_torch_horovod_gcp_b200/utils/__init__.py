"""Utilities: timeline tracing, GPU discovery, clock sampling, throughput meters."""
from .timeline import Timeline  # noqa: F401
from .gpus import getGPUs, GPU  # noqa: F401
