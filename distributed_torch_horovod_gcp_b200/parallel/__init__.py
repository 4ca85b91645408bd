"""Data-parallel machinery: static bucket planner, fused allreduce+optimizer engine,
side-stream overlap."""
from .buckets import Bucket, plan_buckets, plan_hash  # noqa: F401
