"""Comparison baselines — all of them stay on NCCL (BASELINE.json north_star):
WFBP / MG-WFBP / ASC / MGS (wfbp.py), PyTorch-DDP and ZeRO-1 (ddp.py), Horovod-style fusion and
ByteScheduler-style partitioning emulations (horovod_like.py)."""
from .wfbp import DistributedOptimizer as WFBPDistributedOptimizer, mgwfbp_groups, mgs_groups, threshold_groups  # noqa: F401
from .ddp import wrap_ddp  # noqa: F401
