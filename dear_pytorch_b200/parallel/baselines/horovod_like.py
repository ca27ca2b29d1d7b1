"""Horovod- and ByteScheduler-style baselines emulated on torch.distributed (neither library is
installable here; reference drivers: horovod/*_benchmark.py, bytescheduler/imagenet_benchmark.py:73-82).

  * Horovod tensor fusion: gradients that become ready within a fusion window are packed into one
    buffer of at most ``fusion_threshold_mb`` (default 64 MB) and all-reduced together;
  * ByteScheduler: every gradient is partitioned into ``partition_mb`` (4 MB) chunks, chunks are
    all-reduced in *forward-priority* order under a ``credit_mb`` (16 MB) in-flight budget
    (bytescheduler/horovod_mpi_cj.sh:24-27).
Both reuse the WFBP optimizer's machinery with a different grouping.
"""
from .wfbp import DistributedOptimizer as _WFBP


def HorovodLikeOptimizer(optimizer, model, fusion_threshold_mb: float = 64.0, **kw):
    return _WFBP(optimizer, model=model, threshold=int(fusion_threshold_mb * 1024 * 1024 / 4), **kw)


def ByteSchedulerLikeOptimizer(optimizer, model, partition_mb: float = 4.0, credit_mb: float = 16.0, **kw):
    # partitioned groups of ~partition_mb; the credit bounds how many are in flight, which on one
    # NCCL stream is implicit (FIFO), so the grouping is what remains observable
    return _WFBP(optimizer, model=model, threshold=int(partition_mb * 1024 * 1024 / 4), **kw)
