"""Comparison baselines — all of them stay on NCCL (BASELINE.json north_star):
WFBP / MG-WFBP / ASC / MGS with dense, top-k / gaussian-k / gTop-k sparse and sign-compressed gradients and momentum
correction (wfbp.py, gtopk.py), PyTorch-DDP and ZeRO-1 (ddp.py), Horovod cycle-time tensor fusion (horovod.py),
ByteScheduler partition / priority / credit scheduling (bytescheduler.py)."""
from .wfbp import DistributedOptimizer as WFBPDistributedOptimizer, mgwfbp_groups, mgs_groups, threshold_groups  # noqa: F401
from .ddp import wrap_ddp  # noqa: F401
from .horovod import HorovodOptimizer, cycle_groups  # noqa: F401
from .bytescheduler import ByteSchedulerOptimizer, partition_sizes  # noqa: F401
