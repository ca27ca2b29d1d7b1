"""Horovod-style tensor fusion on torch.distributed (Horovod itself is not installable here).

Reference drivers: horovod/imagenet_benchmark.py, horovod/bert_benchmark.py (``hvd.DistributedOptimizer`` with the
library defaults); the knobs the reference's other drivers pin to zero to switch fusion OFF
(``HOROVOD_FUSION_THRESHOLD``, ``HOROVOD_CYCLE_TIME``, ``HOROVOD_CACHE_CAPACITY``;
dear/imagenet_benchmark.py:10-12) are exactly the ones this baseline implements:

  * **cycle time** — Horovod's background loop wakes every ``HOROVOD_CYCLE_TIME`` ms (default 5; this is also the
    ``CYCLE_TIME = 5`` of dear/dopt_rsag_wt.py:40), collects every gradient that became ready on ALL ranks since the
    last cycle and fuses them, in readiness order, into buffers of at most ``HOROVOD_FUSION_THRESHOLD`` bytes
    (default 64 MB), one all-reduce per buffer;
  * **response cache** — the negotiation result is cached, so after the first iterations the same tensors are
    fused into the same groups without another negotiation round.

  * **compression / reduction op** — ``hvd.Compression.fp16`` (``--fp16-allreduce`` of horovod/imagenet_benchmark.py:19,81:
    the fused buffer travels as fp16) and ``op=hvd.Adasum`` (``--use-adasum``, :39,87: scale-insensitive pairwise
    combination  a (+) b = (1 - a.b / 2|a|^2) a + (1 - a.b / 2|b|^2) b  applied per tensor over log2(P) recursive-doubling
    rounds; no averaging afterwards and no lr scaling by the world size).

Emulation: during ``negotiation_steps`` warm-up iterations every gradient is all-reduced on its own (cold cache) while
rank 0 records *when* each gradient hook fired relative to the first one.  Rank 0 then cuts that timeline into
``cycle_time_ms`` windows, splits windows at the fusion threshold, and broadcasts the grouping — the cached
responses.  From then on a fused buffer is all-reduced as soon as its last member's gradient arrives.  Groups are
identical on every rank by construction (the property Horovod's coordinator provides).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import runtime


def _flat_dense(t: torch.Tensor) -> torch.Tensor:
    """1-D alias of a dense tensor's storage (channels-last gradients are dense but not contiguous)."""
    if t.is_contiguous():
        return t.view(-1)
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


def adasum_combine(a: torch.Tensor, b: torch.Tensor, bounds: List[tuple]) -> torch.Tensor:
    """Adasum of two flat buffers, tensor by tensor (``bounds`` = [(start, end), ...] of the member tensors): orthogonal
    gradients add, parallel ones average.  Symmetric in (a, b), so both partners of an exchange compute the same bits."""
    out = torch.empty_like(a)
    for s0, s1 in bounds:
        x, y = a[s0:s1], b[s0:s1]
        xf, yf = x.float(), y.float()
        dot, xx, yy = torch.dot(xf, yf), torch.dot(xf, xf), torch.dot(yf, yf)
        cx = torch.where(xx > 0, 1.0 - dot / (2.0 * xx), torch.ones_like(dot))
        cy = torch.where(yy > 0, 1.0 - dot / (2.0 * yy), torch.ones_like(dot))
        out[s0:s1] = (cx * xf + cy * yf).to(a.dtype)
    return out


def cycle_groups(arrival_ms: List[float], nbytes: List[int], cycle_time_ms: float, fusion_threshold_bytes: int) -> List[List[int]]:
    """Group tensor indices (given in readiness order with their ready times) the way Horovod's cycle does:
    a new group starts at every cycle boundary and whenever the fusion buffer would overflow.  A tensor larger
    than the threshold travels alone.  ``cycle_time_ms <= 0`` or a zero threshold disable fusion (one tensor per
    all-reduce), like the reference's DeAR/WFBP drivers configure Horovod."""
    if not arrival_ms:
        return []
    if cycle_time_ms <= 0 or fusion_threshold_bytes <= 0:
        return [[i] for i in range(len(arrival_ms))]
    groups, cur, cur_bytes = [], [], 0
    t0 = arrival_ms[0]
    cur_cycle = 0
    for i, (t, nb) in enumerate(zip(arrival_ms, nbytes)):
        cyc = int((t - t0) // cycle_time_ms)
        if cur and (cyc != cur_cycle or cur_bytes + nb > fusion_threshold_bytes):
            groups.append(cur)
            cur, cur_bytes = [], 0
        cur_cycle = cyc
        cur.append(i)
        cur_bytes += nb
    if cur:
        groups.append(cur)
    return groups


class _HorovodOptimizer(torch.optim.Optimizer):
    def __init__(self, params, named_parameters, fusion_threshold_mb=None, cycle_time_ms=None, negotiation_steps=2,
                 verbose=True, fp16_allreduce=False, op="average"):
        super(self.__class__, self).__init__(params)
        if op not in ("average", "adasum"):
            raise ValueError("op must be 'average' or 'adasum'")
        self._fp16, self._adasum = bool(fp16_allreduce), op == "adasum"
        if not runtime.is_initialized():
            runtime.init()
        self._rank, self._world, self._device = runtime.rank(), runtime.size(), runtime.device()
        self._pg = runtime.group()
        if self._adasum and self._world & (self._world - 1):
            raise ValueError("Adasum needs a power-of-two number of ranks (recursive doubling), got %d" % self._world)
        if fusion_threshold_mb is None:
            fusion_threshold_mb = float(os.environ.get("HOROVOD_FUSION_THRESHOLD", 64 * 1024 * 1024)) / (1024 * 1024)
        if cycle_time_ms is None:
            cycle_time_ms = float(os.environ.get("HOROVOD_CYCLE_TIME", 5.0))
        self.fusion_threshold_bytes = int(fusion_threshold_mb * 1024 * 1024)
        self.cycle_time_ms = float(cycle_time_ms)
        self._negotiation_steps = max(1, int(negotiation_steps))
        named = [(n, p) for n, p in named_parameters if p.requires_grad]
        self._names = {p: n for n, p in named}
        self._by_name = dict(named)
        self._params = [p for _, p in named]
        self._cuda = self._device.type == "cuda"
        self._stream = torch.cuda.Stream(device=self._device, priority=-1) if self._cuda else None
        self._steps = 0
        self._arrival: List[tuple] = []          # (name, ms since the first hook of this backward)
        self._t_first: Optional[float] = None
        self.groups: Optional[List[List[str]]] = None   # the cached responses
        self._group_of: Dict[str, int] = {}
        self._buffers: Dict[int, torch.Tensor] = {}
        self._offsets: Dict[str, tuple] = {}
        self._bounds: Dict[int, List[tuple]] = {}
        self._arrived: List[int] = []
        self._launched: Dict[object, object] = {}
        self._verbose = verbose and self._rank == 0
        self._hooks = []
        if self._world > 1:
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- comm stream helpers ------------------------------------------------------------------
    def _reduce(self, tensor, bounds):
        """Sum (or Adasum) of ``tensor`` over the ranks, in place; ``bounds`` = member tensors of a fused buffer."""
        wire = tensor.to(torch.float16) if (self._fp16 and tensor.dtype == torch.float32) else tensor
        if not self._adasum:
            dist.all_reduce(wire, group=self._pg)
        else:
            d = 1
            while d < self._world:
                peer = self._rank ^ d
                other = torch.empty_like(wire)
                reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, wire, peer, group=self._pg),
                                               dist.P2POp(dist.irecv, other, peer, group=self._pg)])
                for r in reqs:
                    r.wait()
                lo, hi = (wire, other) if self._rank < peer else (other, wire)
                wire = adasum_combine(lo, hi, bounds)
                d *= 2
        if wire is not tensor:
            tensor.copy_(wire)

    def _allreduce(self, tensor, bounds=None):
        bounds = bounds if bounds is not None else [(0, tensor.numel())]
        if self._cuda:
            self._stream.wait_stream(torch.cuda.current_stream(self._device))
            with torch.cuda.stream(self._stream):
                self._reduce(tensor, bounds)
        else:
            self._reduce(tensor, bounds)

    # ---- backward hook ------------------------------------------------------------------------
    def _on_grad(self, p):
        n = self._names[p]
        if self.groups is None:
            # cold response cache: every tensor is negotiated and reduced on its own
            now = time.perf_counter()
            if self._t_first is None:
                self._t_first = now
            self._arrival.append((n, (now - self._t_first) * 1e3))
            self._allreduce(_flat_dense(p.grad))
            self._launched[n] = True
            return
        gi = self._group_of[n]
        a, b = self._offsets[n]
        self._buffers[gi][a:b].copy_(p.grad.reshape(-1))
        self._arrived[gi] += 1
        if self._arrived[gi] == len(self.groups[gi]):
            self._allreduce(self._buffers[gi], self._bounds[gi])
            self._launched[gi] = True

    def _build_groups(self):
        """Rank 0 turns its recorded readiness timeline into fusion groups; everyone adopts them."""
        plan = None
        if self._rank == 0:
            seen, names, times = set(), [], []
            for n, t in self._arrival:
                if n not in seen:
                    seen.add(n); names.append(n); times.append(t)
            for n in self._by_name:                     # parameters that never produced a gradient: own group each
                if n not in seen:
                    names.append(n); times.append((times[-1] if times else 0.0) + 10 * max(self.cycle_time_ms, 1.0))
            nbytes = [self._by_name[n].numel() * self._by_name[n].element_size() for n in names]
            idx = cycle_groups(times, nbytes, self.cycle_time_ms, self.fusion_threshold_bytes)
            plan = [[names[i] for i in g] for g in idx]
        plan = runtime.broadcast_object(plan, src=0)
        # a fusion buffer holds one dtype
        groups = []
        for g in plan:
            by_dt: Dict[torch.dtype, List[str]] = {}
            for n in g:
                by_dt.setdefault(self._by_name[n].dtype, []).append(n)
            groups.extend(by_dt.values())
        self.groups = groups
        self._group_of = {n: gi for gi, g in enumerate(groups) for n in g}
        for gi, g in enumerate(groups):
            off = 0
            for n in g:
                k = self._by_name[n].numel()
                self._offsets[n] = (off, off + k)
                off += k
            self._buffers[gi] = torch.zeros(off, device=self._device, dtype=self._by_name[g[0]].dtype)
            self._bounds[gi] = [self._offsets[n] for n in g]
        self._arrived = [0] * len(groups)
        if self._verbose:
            sizes = [sum(self._by_name[n].numel() * self._by_name[n].element_size() for n in g) / 2 ** 20 for g in groups]
            print("[horovod-like] response cache built: %d fused all-reduces per step for %d tensors "
                  "(cycle %.1f ms, threshold %.0f MB; largest %.1f MB)" % (
                      len(groups), len(self._params), self.cycle_time_ms, self.fusion_threshold_bytes / 2 ** 20, max(sizes)))

    def synchronize(self):
        if self._world == 1:
            return
        if self.groups is None:
            for p in self._params:                              # gradients that never arrived this step
                if self._names[p] not in self._launched and p.grad is not None:
                    self._allreduce(_flat_dense(p.grad))
            if self._cuda:
                torch.cuda.current_stream(self._device).wait_stream(self._stream)
            if not self._adasum:
                for p in self._params:
                    if p.grad is not None:
                        p.grad.div_(self._world)
            self._launched.clear()
            self._t_first = None
            return
        for gi in range(len(self.groups)):
            if gi not in self._launched:
                self._allreduce(self._buffers[gi], self._bounds[gi])
        if self._cuda:
            torch.cuda.current_stream(self._device).wait_stream(self._stream)
        for gi, g in enumerate(self.groups):
            buf = self._buffers[gi]
            if not self._adasum:
                buf.div_(self._world)
            for n in g:
                a, b = self._offsets[n]
                p = self._by_name[n]
                if p.grad is not None:
                    p.grad.copy_(buf[a:b].view_as(p.grad))
            buf.zero_()                                          # a member that misses a gradient next step adds zeros
            self._arrived[gi] = 0
        self._launched.clear()

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.synchronize()
        super(self.__class__, self).step()
        self._steps += 1
        if self.groups is None and self._world > 1 and self._steps >= self._negotiation_steps:
            self._build_groups()
        self._arrival.clear()
        return loss


def HorovodOptimizer(optimizer, model: Optional[nn.Module] = None, named_parameters=None, fusion_threshold_mb=None,
                     cycle_time_ms=None, negotiation_steps: int = 2, verbose: bool = True, fp16_allreduce: bool = False,
                     op: str = "average", **ignored):
    """``hvd.DistributedOptimizer`` look-alike with cycle-time tensor fusion (see the module docstring).
    ``fp16_allreduce=True`` = ``compression=hvd.Compression.fp16``; ``op="adasum"`` = ``op=hvd.Adasum``."""
    if named_parameters is None:
        if model is None:
            raise ValueError("pass model or named_parameters")
        named_parameters = model.named_parameters()
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_HorovodOptimizer.__dict__))
    return cls(optimizer.param_groups, list(named_parameters), fusion_threshold_mb=fusion_threshold_mb,
               cycle_time_ms=cycle_time_ms, negotiation_steps=negotiation_steps, verbose=verbose,
               fp16_allreduce=fp16_allreduce, op=op)
