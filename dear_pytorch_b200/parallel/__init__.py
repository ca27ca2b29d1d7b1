"""Data-parallel strategies: the DeAR engine, its variants, and the NCCL baselines."""
from .bucket import BucketPlan  # noqa: F401
from .optimizer import DistributedOptimizer, DearEngine  # noqa: F401
