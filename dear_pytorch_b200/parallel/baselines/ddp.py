"""PyTorch-DDP baseline (+ optional ZeRO-1), as pytorch-ddp/imagenet_benchmark.py:65-70 of the reference."""
import torch

from ... import runtime


def wrap_ddp(model, optimizer_cls=None, optimizer_kwargs=None, zero: bool = False, bucket_cap_mb: float = 25.0,
             gradient_as_bucket_view: bool = True):
    """Return ``(ddp_model, optimizer)``; the optimizer is a ZeroRedundancyOptimizer when ``zero``."""
    if not runtime.is_initialized():
        runtime.init(backend="nccl" if torch.cuda.is_available() else "gloo")
    dev = runtime.device()
    kwargs = dict(bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=gradient_as_bucket_view)
    if runtime.size() > 1:
        if dev.type == "cuda":
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], **kwargs)
        else:
            model = torch.nn.parallel.DistributedDataParallel(model, **kwargs)
    opt = None
    if optimizer_cls is not None:
        optimizer_kwargs = optimizer_kwargs or {}
        if zero and runtime.size() > 1:
            from torch.distributed.optim import ZeroRedundancyOptimizer
            opt = ZeroRedundancyOptimizer(model.parameters(), optimizer_class=optimizer_cls, **optimizer_kwargs)
        else:
            opt = optimizer_cls(model.parameters(), **optimizer_kwargs)
    return model, opt
