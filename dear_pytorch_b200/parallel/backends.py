"""Data-path backends of the DeAR engine.

``NativeBackend``  (b200 / emu): symmetric buckets + the two fused kernels of
                   csrc/kernels.cu (or their host emulation) — no NCCL call and no separate
                   elementwise kernel on either the backward or the forward path.
``TorchBackend``   (nccl / gloo): the same bucket layout driven by
                   ``torch.distributed.reduce_scatter_tensor`` / ``all_gather_into_tensor`` and
                   an eager sharded SGD — the comparison baseline and the CPU plumbing path.
                   This is what the reference does per bucket (dear/tensorfusion.py:469-482,
                   dear/dear_dopt.py:293-336), minus its per-parameter loops.

Both expose the same interface to ``parallel.optimizer``:
  param_buffer(g) / grad_buffer(g)      flat bucket tensors (parameters / gradients are views)
  reduce_scatter(g, pack)               backward-path collective (+ 1/P scale)
  allgather_update(g, ...)              forward-path collective (+ sharded SGD)
  wait_bucket(g) / wait_all() / fence() / synchronize()
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops
from .bucket import BucketPlan

_DT_CODE = {torch.float32: "DT_F32", torch.bfloat16: "DT_BF16", torch.float16: "DT_F16"}


OPT_SGD, OPT_ADAM, OPT_ADAMW = 0, 1, 2
HYPER_NESTEROV, HYPER_SKIP = 1, 2        # bits of the ``nesterov`` field of a hyper segment (csrc/dear_common.h)


class HyperSpec:
    """Per-bucket hyper-parameter segments:
    ``[(end_elem, lr, wd, momentum|beta1, dampening, nesterov, opt, beta2, eps)]`` (opt: 0 SGD, 1 Adam, 2 AdamW)."""

    __slots__ = ("segs",)

    def __init__(self, segs):
        self.segs = tuple(tuple(s) + (OPT_SGD, 0.0, 0.0)[len(s) - 6:] if len(s) < 9 else tuple(s) for s in segs)

    def __eq__(self, other):
        return isinstance(other, HyperSpec) and self.segs == other.segs

    @property
    def is_adam(self):
        return any(s[6] != OPT_SGD for s in self.segs)

    @property
    def uses_momentum(self):
        return self.is_adam or any(s[3] > 0 for s in self.segs)


class _BackendBase:
    steal_grads = False

    def __init__(self, plan: BucketPlan, rank: int, world: int, device: torch.device):
        self.plan = plan
        self.rank = rank
        self.world = world
        self.device = device
        nb = len(plan.buckets)
        self.grad_shard: List[torch.Tensor] = [None] * nb
        self.mom_shard: List[Optional[torch.Tensor]] = [None] * nb
        self.master_shard: List[Optional[torch.Tensor]] = [None] * nb
        self.var_shard: List[Optional[torch.Tensor]] = [None] * nb       # Adam exp_avg_sq
        self.hyper: List[Optional[HyperSpec]] = [None] * nb
        self.grad_scale = 1.0

    # -- shard state ------------------------------------------------------------------
    def _alloc_shards(self):
        for b in self.plan.buckets:
            self.grad_shard[b.index] = torch.zeros(b.shard_numel, dtype=torch.float32, device=self.device)

    def init_master_shards(self):
        """fp32 master copy of this rank's shard for low-precision parameter buckets."""
        for b in self.plan.buckets:
            if b.dtype != torch.float32:
                lo = self.rank * b.shard_numel
                self.master_shard[b.index] = self.param_buffer(b.index)[lo:lo + b.shard_numel].float().clone()
        self._shards_changed()

    def ensure_momentum(self, g: int):
        if self.mom_shard[g] is None:
            self.mom_shard[g] = torch.zeros(self.plan.buckets[g].shard_numel, dtype=torch.float32, device=self.device)
            self._shards_changed(g)

    def ensure_var(self, g: int):
        if self.var_shard[g] is None:
            self.var_shard[g] = torch.zeros(self.plan.buckets[g].shard_numel, dtype=torch.float32, device=self.device)
            self._shards_changed(g)

    def _shards_changed(self, g: Optional[int] = None):
        pass

    def set_step(self, t: int) -> None:
        """Number of updates already applied (Adam bias correction); called after (re)building buckets."""
        pass

    def set_hyper(self, g: int, spec: HyperSpec) -> None:
        self.hyper[g] = spec

    def launches(self) -> int:
        return 0

    def set_grad_scale(self, s: float) -> None:
        """Extra factor applied to the averaged gradient (1/loss_scale for static loss scaling)."""
        self.grad_scale = float(s)

    def stream_key(self, g: int) -> int:
        """Buckets with the same key share one communication stream (ordering domain)."""
        return 0


# =====================================================================================
# native: b200 kernels / host emulation
# =====================================================================================
class NativeBackend(_BackendBase):
    steal_grads = True

    def __init__(self, comm, plan: BucketPlan, rank: int, world: int, device: torch.device):
        super().__init__(plan, rank, world, device)
        C = ops.require_native()
        self.C = C
        self.comm = comm
        # one native BucketSet per dtype; bucket g -> (set, local index)
        by_dtype: Dict[torch.dtype, List[int]] = {}
        for b in plan.buckets:
            if b.dtype not in _DT_CODE:
                raise TypeError("unsupported parameter dtype %s" % b.dtype)
            by_dtype.setdefault(b.dtype, []).append(b.index)
        self.sets = {}
        self.where: List[Tuple[object, int]] = [None] * len(plan.buckets)
        for dt, idxs in by_dtype.items():
            bs = C.BucketSet(comm, [plan.buckets[g].padded_numel for g in idxs], getattr(C, _DT_CODE[dt]), True)
            self.sets[dt] = bs
            for li, g in enumerate(idxs):
                self.where[g] = (bs, li)
        self._pbuf = [bs.param_buffer(li) for bs, li in self.where]
        self._gbuf = [bs.grad_buffer(li) for bs, li in self.where]
        self._alloc_shards()
        self._shards_changed()
        self._first_in_set = {id(bs): min(g for g, (s, _) in enumerate(self.where) if s is bs)
                              for bs in self.sets.values()}

    @property
    def has_multicast(self) -> bool:
        return any(bs.has_multicast() for bs in self.sets.values())

    def param_buffer(self, g):
        return self._pbuf[g]

    def grad_buffer(self, g):
        return self._gbuf[g]

    def _shards_changed(self, g=None):
        for i in (range(len(self.where)) if g is None else (g,)):
            bs, li = self.where[i]
            if self.grad_shard[i] is not None:
                bs.set_shards(li, self.grad_shard[i], self.mom_shard[i], self.master_shard[i], self.var_shard[i])

    def set_hyper(self, g, spec: HyperSpec):
        if self.hyper[g] == spec:
            return
        self.hyper[g] = spec
        if spec.uses_momentum:
            self.ensure_momentum(g)
        if spec.is_adam:
            self.ensure_var(g)
        bs, li = self.where[g]
        bs.set_hyper(li, [s[0] for s in spec.segs], [s[1] for s in spec.segs], [s[2] for s in spec.segs],
                     [s[3] for s in spec.segs], [s[4] for s in spec.segs], [int(s[5]) for s in spec.segs],
                     [int(s[6]) for s in spec.segs], [s[7] for s in spec.segs], [s[8] for s in spec.segs])

    def set_step(self, t: int):
        for i, (bs, li) in enumerate(self.where):
            bs.set_step(li, int(t))

    def set_pack(self, g, src_ptrs, dst_off, nbytes, flags):
        bs, li = self.where[g]
        bs.set_pack(li, src_ptrs, dst_off, nbytes, flags)

    def reduce_scatter(self, g, pack=True):
        bs, li = self.where[g]
        bs.reduce_scatter(li, pack)

    def allgather_update(self, g, do_update=True, first_step=False, zero_grad=False):
        bs, li = self.where[g]
        # the first bucket of every set carries the entry rendezvous: nobody overwrites a
        # peer's parameters before that peer has finished its backward pass.
        bs.allgather_update(li, do_update, first_step, self._first_in_set[id(bs)] == g, zero_grad)

    def fence(self):
        for bs in self.sets.values():
            bs.fence_current_to_comm()

    def wait_bucket(self, g):
        bs, li = self.where[g]
        bs.wait_bucket(li)

    def wait_all(self):
        for bs in self.sets.values():
            bs.wait_all()

    def synchronize(self):
        for bs in self.sets.values():
            bs.synchronize()

    def launches(self):
        return self.comm.launches()

    def set_grad_scale(self, s):
        self.grad_scale = float(s)
        for bs in self.sets.values():
            bs.set_grad_scale(float(s))

    def stream_key(self, g):
        return id(self.where[g][0])


# =====================================================================================
# torch.distributed: nccl / gloo
# =====================================================================================
class TorchBackend(_BackendBase):
    steal_grads = False

    def __init__(self, group, plan: BucketPlan, rank: int, world: int, device: torch.device):
        super().__init__(plan, rank, world, device)
        self.group = group
        self.cuda = device.type == "cuda"
        self._pbuf = [torch.zeros(b.padded_numel, dtype=b.dtype, device=device) for b in plan.buckets]
        self._gbuf = [torch.zeros(b.padded_numel, dtype=b.dtype, device=device) for b in plan.buckets]
        self._rs_out = [torch.zeros(b.shard_numel, dtype=b.dtype, device=device) for b in plan.buckets]
        self._alloc_shards()
        self._n_launch = 0
        self._t = 0                      # updates applied so far (Adam bias correction)
        if self.cuda:
            self.stream = torch.cuda.Stream(device=device, priority=-1)
            self.ag_done = [torch.cuda.Event() for _ in plan.buckets]
            self._pending = [False] * len(plan.buckets)

    def param_buffer(self, g):
        return self._pbuf[g]

    def grad_buffer(self, g):
        return self._gbuf[g]

    def set_hyper(self, g, spec: HyperSpec):
        self.hyper[g] = spec
        if spec.uses_momentum:
            self.ensure_momentum(g)
        if spec.is_adam:
            self.ensure_var(g)

    def set_step(self, t: int):
        self._t = int(t)

    def set_pack(self, g, src_ptrs, dst_off, nbytes, flags):
        pass    # gradients are accumulated straight into the bucket views

    def _on_comm_stream(self):
        if self.cuda:
            return torch.cuda.stream(self.stream)
        import contextlib
        return contextlib.nullcontext()

    def reduce_scatter(self, g, pack=True):
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with self._on_comm_stream():
            if self.world > 1:
                dist.reduce_scatter_tensor(self._rs_out[g], self._gbuf[g], op=dist.ReduceOp.SUM, group=self.group)
            else:
                self._rs_out[g].copy_(self._gbuf[g])
            torch.mul(self._rs_out[g].float(), self.grad_scale / self.world, out=self.grad_shard[g])
            self._gbuf[g].zero_()
        self._n_launch += 3

    @torch.no_grad()
    def _sgd_shard(self, g, first_step):
        b = self.plan.buckets[g]
        lo, hi = self.rank * b.shard_numel, (self.rank + 1) * b.shard_numel
        master = self.master_shard[g]
        start = 0
        for (end, lr, wd, mom, damp, nesterov, opt, beta2, eps) in self.hyper[g].segs:
            a, z = max(start, lo), min(end, hi)
            start = end
            if a >= z:
                continue
            if int(nesterov) & HYPER_SKIP and not bool(self.grad_shard[g][a - lo:z - lo].any()):
                continue                          # no gradient on ANY rank this step: parameter and state stay as they are
            nesterov = bool(int(nesterov) & HYPER_NESTEROV)
            sl = slice(a - lo, z - lo)
            p = master[sl] if master is not None else self._pbuf[g][a:z]
            d = self.grad_shard[g][sl]
            if opt != OPT_SGD:
                t = self._t + 1
                m, v = self.mom_shard[g][sl], self.var_shard[g][sl]
                if opt == OPT_ADAM and wd != 0:
                    d = d.add(p, alpha=wd)
                m.mul_(mom).add_(d, alpha=1 - mom)
                v.mul_(beta2).addcmul_(d, d, value=1 - beta2)
                bc1, bc2 = 1 - mom ** t, 1 - beta2 ** t
                if opt == OPT_ADAMW:
                    p.mul_(1 - lr * wd)
                p.addcdiv_(m, v.sqrt().div_(bc2 ** 0.5).add_(eps), value=-lr / bc1)
                self._n_launch += 6
                continue
            if wd != 0:
                d = d.add(p, alpha=wd)
            if mom > 0:
                buf = self.mom_shard[g][sl]
                if first_step:
                    buf.copy_(d)
                else:
                    buf.mul_(mom).add_(d, alpha=1 - damp)
                d = d.add(buf, alpha=mom) if nesterov else buf
            p.add_(d, alpha=-lr)
            self._n_launch += 3

    def allgather_update(self, g, do_update=True, first_step=False, zero_grad=False):
        b = self.plan.buckets[g]
        lo = self.rank * b.shard_numel
        with self._on_comm_stream():
            if do_update:
                self._sgd_shard(g, first_step)
            if self.master_shard[g] is not None:
                src = self.master_shard[g].to(b.dtype)
            else:
                src = self._pbuf[g][lo:lo + b.shard_numel].clone()
            if self.world > 1:
                dist.all_gather_into_tensor(self._pbuf[g], src, group=self.group)
            else:
                self._pbuf[g][lo:lo + b.shard_numel].copy_(src)
            if do_update and g == len(self.plan.buckets) - 1:
                self._t += 1
            if self.cuda:
                self.ag_done[g].record(self.stream)
                self._pending[g] = True
        self._n_launch += 2

    def fence(self):
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def wait_bucket(self, g):
        if self.cuda and self._pending[g]:
            torch.cuda.current_stream(self.device).wait_event(self.ag_done[g])

    def wait_all(self):
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def synchronize(self):
        if self.cuda:
            self.stream.synchronize()

    def launches(self):
        return self._n_launch
