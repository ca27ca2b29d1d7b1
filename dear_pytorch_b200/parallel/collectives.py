"""Tensor-level collectives of the public API.

``allreduce``               dear/dear_dopt.py:546-549  (sum, then divide by size)
``broadcast_parameters``    dear/dear_dopt.py:400-425
``broadcast_optimizer_state`` dear/dear_dopt.py:428-544 (Horovod-derived; broken in the reference:
                            ``collections.Iterable`` and CPU tensors through NCCL — re-done here)
On the b200/emu backends these run our general-purpose kernel (csrc/kernels.cu: gen_kernel) on the
symmetric staging buffer; on nccl/gloo they are torch.distributed calls.
"""
from __future__ import annotations

import collections.abc

import torch
import torch.distributed as dist

from .. import runtime


def _flat_dense(t: torch.Tensor) -> torch.Tensor:
    """1-D alias of a dense tensor's storage range (no copy), whatever its memory format."""
    if t.is_contiguous():
        return t.view(-1) if t.dim() else t.reshape(1)
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


def _is_dense(t: torch.Tensor) -> bool:
    if t.is_contiguous():
        return True
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return True
    if t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d):
        return True
    return False


def _staged(t: torch.Tensor):
    """Return (tensor usable by the data path, needs_copy_back)."""
    dev = runtime.device()
    if t.device.type != dev.type or not _is_dense(t):
        return t.detach().to(dev).contiguous(), True
    return t.detach(), False


def allreduce_(tensor: torch.Tensor, average: bool = True) -> torch.Tensor:
    """In-place all-reduce; stream-ordered (no host block) on GPU backends."""
    world = runtime.size()
    if world == 1:
        return tensor
    work, back = _staged(tensor)
    is_float = work.dtype in (torch.float32, torch.bfloat16, torch.float16)
    comm = runtime.communicator()
    if comm is not None:
        buf = _flat_dense(work) if is_float else work.reshape(-1).float()
        h = comm.allReduce(buf, (1.0 / world) if average else 1.0)
        comm.waitStream(h)
        if not is_float:
            work.reshape(-1).copy_(buf)
    else:
        dist.all_reduce(work, op=dist.ReduceOp.SUM, group=runtime.group())
        if average:
            if is_float:
                work.div_(world)
            else:
                work.copy_(work.float().div_(world))
    if back:
        tensor.copy_(work.reshape(tensor.shape))
    return tensor


def allreduce(tensor: torch.Tensor, name=None) -> torch.Tensor:
    """Reference semantics: in-place sum all-reduce, return ``tensor / size()``
    (dear/dear_dopt.py:546-549).  The division is fused into the kernel here."""
    return allreduce_(tensor, average=True)


def broadcast_(tensor: torch.Tensor, root_rank: int = 0) -> torch.Tensor:
    if runtime.size() == 1:
        return tensor
    work, back = _staged(tensor)
    comm = runtime.communicator()
    if comm is not None:
        h = comm.bcast(_flat_dense(work), root_rank)
        comm.waitStream(h)
    else:
        dist.broadcast(work, src=root_rank, group=runtime.group())
    if back:
        tensor.copy_(work.reshape(tensor.shape))
    return tensor


def broadcast_parameters(params, root_rank: int = 0) -> None:
    """Broadcast a ``state_dict()`` (or an iterable of ``(name, tensor)``) from ``root_rank``."""
    if isinstance(params, dict):
        items = sorted(params.items())
    elif isinstance(params, collections.abc.Iterable):
        items = list(params)
        if items and not isinstance(items[0], tuple):
            items = [("param.noname.%d" % i, p) for i, p in enumerate(items)]
    else:
        raise ValueError("invalid params of type: %s" % type(params))
    for _, p in items:
        if p is None or not torch.is_tensor(p):
            continue
        broadcast_(p, root_rank)
    comm = runtime.communicator()
    if comm is not None:
        comm.synchronize()
    elif runtime.device().type == "cuda":
        torch.cuda.current_stream().synchronize()
    # parameters that already live in DeAR buckets: their sharded fp32 masters were snapshotted when the optimizer
    # was wrapped and must follow the broadcast values (bf16 / fp16 models)
    from .optimizer import live_engines
    for eng in live_engines():
        eng.params_changed()


def broadcast_optimizer_state(optimizer: torch.optim.Optimizer, root_rank: int = 0) -> None:
    """Make every rank's optimizer hyper-parameters and state equal to ``root_rank``'s."""
    if runtime.size() == 1:
        return
    eng = getattr(optimizer, "_dear", None)
    # (1) scalar options of every param group, as one pickled object
    opts = [{k: v for k, v in g.items() if k != "params"} for g in optimizer.param_groups]
    opts = runtime.broadcast_object(opts, src=root_rank)
    for g, o in zip(optimizer.param_groups, opts):
        g.update(o)
    # (2) tensor state
    if eng is not None:
        eng.synchronize(host=True)
        init = runtime.broadcast_object(eng._mom_initialised, src=root_rank)
        eng._mom_initialised = bool(init)
        # sharded state: shard r is only meaningful on rank r and was derived from identical
        # (broadcast) parameters, so nothing to move; master shards follow the parameters.
        eng.backend.init_master_shards()
        return
    for group in optimizer.param_groups:
        for p in group["params"]:
            st = optimizer.state.get(p, {})
            # the root decides which entries exist and what they look like: a rank that has not stepped yet
            # (fresh start next to a resumed root) has no state and must allocate before it can receive
            meta = None
            if runtime.rank() == root_rank:
                meta = [(k, (tuple(v.shape), v.dtype) if torch.is_tensor(v) else None) for k, v in sorted(st.items())]
            meta = runtime.broadcast_object(meta, src=root_rank)
            for k, tinfo in meta:
                if tinfo is not None:
                    v = st.get(k)
                    if not torch.is_tensor(v) or tuple(v.shape) != tinfo[0] or v.dtype != tinfo[1]:
                        v = torch.zeros(tinfo[0], dtype=tinfo[1], device=p.device)
                        st[k] = v
                    broadcast_(v, root_rank)
                else:
                    st[k] = runtime.broadcast_object(st.get(k), src=root_rank)
            if meta and p not in optimizer.state:
                optimizer.state[p] = st


def allgather(tensor: torch.Tensor) -> torch.Tensor:
    """Concatenate ``tensor`` from every rank along dim 0 (equal sizes)."""
    world = runtime.size()
    if world == 1:
        return tensor.clone()
    work, _ = _staged(tensor)
    out = torch.empty((world * work.numel(),), dtype=work.dtype, device=work.device)
    comm = runtime.communicator()
    if comm is not None:
        h = comm.allGather(_flat_dense(work), out)
        comm.waitStream(h)
    else:
        dist.all_gather_into_tensor(out, work.reshape(-1), group=runtime.group())
    return out.reshape((world * tensor.shape[0],) + tuple(tensor.shape[1:])) if tensor.dim() else out
