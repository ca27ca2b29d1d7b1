"""DeAR variants shipped by the reference as separate modules.

  naive per-tensor RS/AG ("DeAR without tensor fusion")   dear/dopt_rsag_naive.py
  wait-time-driven bucketing (experimental)                dear/dopt_rsag_wt.py
  reduce / broadcast decoupling (experimental)             dear/dopt_rb.py
  Bayesian-tuned buckets                                   dear/dopt_rsag_bo.py  (see tuner.py)

Here they are options of the same engine instead of ~600-line copies of the optimizer.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch

from .. import runtime
from .bucket import BucketPlan
from .comm import Comm
from .optimizer import DistributedOptimizer

CYCLE_TIME_MS = 5.0       # dear/dopt_rsag_wt.py:40-42
WT_WARMUP_STEPS = 5


# =====================================================================================
# naive: one bucket per module, no fusion
# =====================================================================================
def NaiveDistributedOptimizer(optimizer, model, exclude_parts="", verbose=True, **ignored):
    """Per-module reduce-scatter in backward / all-gather in forward, no tensor fusion.
    Slow by design — start-up cost per tensor (reference docstring dear/dopt_rsag_naive.py:17-19)."""
    excl = exclude_parts.replace("reduce", "reducescatter").replace("bcast", "allgather") \
        if ("reducescatter" not in exclude_parts and "allgather" not in exclude_parts) else exclude_parts
    return DistributedOptimizer(optimizer, model, policy=("per_module",), exclude_parts=excl, verbose=verbose)


# =====================================================================================
# wait-time bucketing
# =====================================================================================
class WaitTimeBucketing:
    """Start with ONE all-layer bucket, measure how long each parameter's gradient waits in the
    buffer (EMA, alpha = 0.9), then cut the model into buckets of ~``cycle_time_ms`` of backward
    time each (dear/dopt_rsag_wt.py:152-192,355-386).  Rank 0's flags are broadcast."""

    def __init__(self, engine, cycle_time_ms: float = CYCLE_TIME_MS, warmup_steps: int = WT_WARMUP_STEPS, alpha: float = 0.9):
        self.eng = engine
        self.cycle = float(cycle_time_ms)
        self.warmup = int(warmup_steps)
        self.alpha = alpha
        self.cuda = engine.device.type == "cuda"
        self.wait_ms: Dict[str, float] = {s.name: 0.0 for s in engine.plan.slots}
        self._in: Dict[str, object] = {}
        self._pending: List = []
        self.done = False
        self.flags: Optional[List[int]] = None
        engine._wt = self
        engine._step_callbacks.append(self._on_step)

    def _now(self):
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return time.perf_counter()

    def param_in(self, slot):
        if not self.done:
            self._in[slot.name] = self._now()

    def bucket_out(self, g):
        if self.done:
            return
        out = self._now()
        for s in self.eng.plan.buckets[g].slots:
            if s.name in self._in:
                self._pending.append((s.name, self._in.pop(s.name), out))

    def _resolve(self):
        for name, t_in, t_out in self._pending:
            if self.cuda:
                t_out.synchronize()
                ms = t_in.elapsed_time(t_out)
            else:
                ms = (t_out - t_in) * 1e3
            self.wait_ms[name] = (1 - self.alpha) * self.wait_ms[name] + self.alpha * ms
        self._pending = []

    def module_wait_times(self) -> List[float]:
        plan = self.eng.plan
        return [max(self.wait_ms[s.name] for s in plan.module_params[mi]) for mi in range(len(plan.modules))]

    def compute_flags(self) -> List[int]:
        """flags[i] == 1 closes a bucket AFTER module i (forward order)."""
        wts = self.module_wait_times()
        cap = self.cycle
        starts = []
        for i, wt in enumerate(wts):
            if i == 0:
                starts.append(1)
            elif wt > cap:
                starts.append(1)
                cap += self.cycle
            else:
                starts.append(0)
        n = len(starts)
        return [1 if (i + 1 < n and starts[i + 1]) or i + 1 == n else 0 for i in range(n)]

    def _on_step(self):
        if self.done:
            return
        self._resolve()
        if self.eng.num_steps == self.warmup:
            flags = self.compute_flags() if self.eng.rank == 0 else None
            flags = runtime.broadcast_object(flags, src=0)
            self.flags = flags
            self.done = True
            self.eng._wt = None
            self.eng.request_rebucket(("flags", tuple(flags)))


def WaitTimeDistributedOptimizer(optimizer, model, cycle_time_ms: float = CYCLE_TIME_MS, warmup_steps: int = WT_WARMUP_STEPS,
                                 exclude_parts="", verbose=True, **ignored):
    opt = DistributedOptimizer(optimizer, model, threshold=None, num_nearby_layers=-1, exclude_parts=exclude_parts,
                               verbose=verbose)
    opt.wait_time = WaitTimeBucketing(opt.engine, cycle_time_ms, warmup_steps)
    return opt


# =====================================================================================
# reduce / broadcast decoupling
# =====================================================================================
class _ReduceBroadcastOptimizer(torch.optim.Optimizer):
    """Gradients are *reduced to rank 0* per bucket during backward; rank 0 applies the update and
    *broadcasts* the parameters bucket by bucket during the next forward (dear/dopt_rb.py:222-294,
    336-372).  Root is fixed to 0 like the reference (:242,301)."""

    ROOT = 0

    def __init__(self, params, model, threshold=25.0, nstreams=2, exclude_parts="", verbose=True):
        super(self.__class__, self).__init__(params)
        if not runtime.is_initialized():
            runtime.init()
        self._rank, self._world, self._device = runtime.rank(), runtime.size(), runtime.device()
        self._comm = Comm(nstreams)
        self._exclude_reduce = "reduce" in exclude_parts
        self._exclude_bcast = "bcast" in exclude_parts
        self._plan = BucketPlan(model, 1).group_by_threshold(threshold)
        self._pbuf, self._gbuf = [], []
        with torch.no_grad():
            for b in self._plan.buckets:
                pb = torch.zeros(b.padded_numel, dtype=b.dtype, device=self._device)
                gb = torch.zeros_like(pb)
                for s in b.slots:
                    p = s.param
                    pv = torch.as_strided(pb, p.shape, p.stride(), s.start) if p.is_contiguous() else None
                    if pv is None:
                        p.data = p.data.contiguous()
                        pv = pb[s.start:s.end].view(p.shape)
                    pv.copy_(p.data)
                    p.data = pv
                    p.grad = torch.as_strided(gb, p.shape, p.stride(), s.start)
                self._pbuf.append(pb)
                self._gbuf.append(gb)
        nb = len(self._plan.buckets)
        self._arrived = [0] * nb
        self._got = set()                  # parameters that received a gradient in the current step
        self._reduce_handle = [None] * nb
        self._bcast_handle = [None] * nb
        self._hooks = []
        for s in self._plan.slots:             # (single process too: the hook records which parameters got a gradient)
            self._hooks.append(s.param.register_post_accumulate_grad_hook(self._on_grad))
        if self._world > 1:
            for mi, m in enumerate(self._plan.modules):
                self._hooks.append(m.register_forward_pre_hook(self._make_pre_hook(mi)))
        if verbose and self._rank == 0:
            print(self._plan.describe())

    def _on_grad(self, p):
        s = self._plan.slot_of[p]
        g = s.bucket
        gv = torch.as_strided(self._gbuf[g], p.shape, p.stride(), s.start)
        if p.grad.data_ptr() != gv.data_ptr():
            gv.copy_(p.grad)
            p.grad = gv
        self._arrived[g] += 1
        self._got.add(p)
        if self._world == 1:
            return
        if self._arrived[g] == len(self._plan.buckets[g].slots) and not self._exclude_reduce:
            self._reduce_handle[g] = self._comm.reduce(self._gbuf[g], self.ROOT, 1.0 / self._world)

    def _make_pre_hook(self, mi):
        def hook(module, inputs):
            g = self._plan.module_bucket[mi]
            h = self._bcast_handle[g]
            if h is not None:
                self._comm.waitStream(h)
                self._bcast_handle[g] = None
        return hook

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._world == 1:
            hidden = [(sl.param, sl.param.grad) for sl in self._plan.slots if sl.param not in self._got]
            for q, _ in hidden:                         # no gradient this step: torch.optim leaves the parameter alone
                q.grad = None
            super(self.__class__, self).step()
            for q, gview in hidden:
                q.grad = gview
            for g, gb in enumerate(self._gbuf):         # gradients are bucket views and zero_grad() is a no-op
                gb.zero_()
                self._arrived[g] = 0
            self._got.clear()
            return loss
        for g, b in enumerate(self._plan.buckets):
            if self._reduce_handle[g] is None and not self._exclude_reduce:
                self._reduce_handle[g] = self._comm.reduce(self._gbuf[g], self.ROOT, 1.0 / self._world)
        for g, b in enumerate(self._plan.buckets):
            if self._reduce_handle[g] is not None:
                self._comm.waitStream(self._reduce_handle[g])
                self._reduce_handle[g] = None
        if self._rank == self.ROOT:
            # a parameter whose gradient is a bucket view always "has" a gradient; torch.optim skips parameters WITHOUT
            # one (no weight decay, no momentum decay).  Hide the views of parameters that received nothing here and
            # whose reduced gradient is zero as well (= unused on every rank) for the duration of the update.
            hidden = []
            if len(self._got) != len(self._plan.slots):
                for sl in self._plan.slots:
                    q = sl.param
                    if q not in self._got and q.grad is not None and not bool(q.grad.any()):
                        hidden.append((q, q.grad))
                        q.grad = None
            super(self.__class__, self).step()          # the user's optimizer, on the averaged gradients
            for q, gview in hidden:
                q.grad = gview
        self._got.clear()
        for g in range(len(self._plan.buckets)):
            self._gbuf[g].zero_()
            self._arrived[g] = 0
            if not self._exclude_bcast:
                self._bcast_handle[g] = self._comm.bcast(self._pbuf[g], self.ROOT)
        return loss

    def zero_grad(self, set_to_none: bool = False):
        return None

    def synchronize(self):
        for g, h in enumerate(self._bcast_handle):
            if h is not None:
                self._comm.waitStream(h)
                self._bcast_handle[g] = None
        self._comm.synchronize()


def ReduceBroadcastDistributedOptimizer(optimizer, model, threshold=25.0, nstreams=2, exclude_parts="", verbose=True, **ignored):
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_ReduceBroadcastOptimizer.__dict__))
    return cls(optimizer.param_groups, model, threshold=threshold, nstreams=nstreams, exclude_parts=exclude_parts,
               verbose=verbose)
