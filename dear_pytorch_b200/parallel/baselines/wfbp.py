"""WFBP / MG-WFBP baselines: gradient all-reduce overlapped with back-propagation.

Reference: wfbp/dopt.py (on comm_core) and {dear,mgwfbp,wfbp}/hv_distributed_optimizer.py (on
Horovod).  Behaviour kept:
  * grouping in REVERSE layer order by number of elements, ``threshold=0`` => per tensor
    (wfbp/dopt.py:321-355);
  * MG-WFBP merging from layer-wise backward times and an alpha-beta all-reduce model
    (wfbp/dopt.py:380-486), ASC variant (hv_distributed_optimizer.py:353-427);
  * dense path = one all-reduce per group launched from the gradient hook; ``step()`` =
    synchronise, average, then the wrapped optimizer's own ``step()`` (wfbp/dopt.py:694-701,807-968);
  * sparse path = compress + all-gather of (values, indices) (wfbp/dopt.py:703-742), or — for the ``gtopk*``
    compressors — the log2(P)-round gTop-k exchange over ``Comm.sendrecv`` (wfbp/dopt.py:50-106,725-728);
  * momentum correction for sparsified training (``momentum_correction=True``; wfbp/dopt.py:769-775,906-953): the
    velocity ``u = m*u + g`` is accumulated locally BEFORE sparsification and is what gets compressed and
    communicated; ``step()`` then applies ``p -= lr*(avg + wd*p)`` without another momentum pass and masks the
    velocity where it was sent (momentum-factor masking via the compressor's ``zero_conditions``).
Deliberate differences: the comm stream waits on the compute stream with an event (the reference
synchronises the host inside the hook, wfbp/dopt.py:696); collectives are torch.distributed NCCL
(comm_core / Horovod are not installable here).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import runtime
from ...utils import perf_model
from ..compression import NoneCompressor, compressors


def threshold_groups(sizes_rev: Sequence[int], threshold: int) -> List[List[int]]:
    """``sizes_rev`` in backward order; close a group once it holds >= ``threshold`` elements."""
    groups, cur, acc = [], [], 0
    for i, n in enumerate(sizes_rev):
        cur.append(i)
        acc += n
        if acc >= threshold:
            groups.append(cur)
            cur, acc = [], 0
    if cur:
        groups.append(cur)
    return groups


def mgwfbp_groups(sizes: Sequence[int], tb: Sequence[float], alpha: float, beta: float, nbytes: int = 4,
                  asc: bool = False, small: int = 8192) -> List[List[int]]:
    """MG-WFBP merged-gradient grouping (Shi et al., INFOCOM 2019).

    ``sizes`` / ``tb``: per-layer element counts and backward times in FORWARD order.  Returns
    groups of layer indices in backward order (first group = last layers).
    """
    L = len(sizes)
    p = list(sizes)
    tc = [perf_model.predict_allreduce_time_with_size(alpha, beta, s * nbytes) for s in p]
    taob = [0.0] * L
    for l in range(L - 2, -1, -1):
        taob[l] = taob[l + 1] + tb[l + 1]

    def comm_start():
        taoc = [0.0] * L
        taoc[L - 1] = taob[L - 1] + tb[L - 1]
        for l in range(L - 2, -1, -1):
            taoc[l] = max(taoc[l + 1] + tc[l + 1], taob[l] + tb[l])
        return taoc

    def merge(l):
        tc[l] = 0.0
        p[l - 1] += p[l]
        p[l] = 0
        tc[l - 1] = perf_model.predict_allreduce_time_with_size(alpha, beta, p[l - 1] * nbytes)

    taoc = comm_start()
    groups, group = [], []
    for l in range(L - 1, 0, -1):
        group.append(l)
        ready_prev = taob[l - 1] + tb[l - 1]          # when layer l-1's gradient is ready
        merged = False
        if ready_prev < taoc[l] + tc[l]:              # comm of l would still be running
            if taoc[l] > ready_prev:                  # ... and has not even started: merging is free
                merge(l); taoc = comm_start(); merged = True
            elif not asc and (ready_prev - taoc[l]) < alpha:   # waiting costs less than a start-up
                merge(l); taoc = comm_start(); merged = True
        if not merged and not asc and p[l] < small:
            merge(l); taoc = comm_start(); merged = True
        if not merged:
            groups.append(group)
            group = []
    group.append(0)
    groups.append(group)
    return groups


def mgs_groups(sizes: Sequence[int], tb: Sequence[float], world: int, density: float) -> List[List[int]]:
    """MGS-SGD merged *sparsified*-gradient grouping (Shi et al., INFOCOM 2020; reference
    wfbp/dopt.py:488-569): merge layer l into l-1 when the extra waiting (longer backward + top-k of
    the merged tensor) is smaller than the all-gather start-up time it saves.  Uses the reference's
    cost models (``perf_model.topk_perf_model`` / ``allgather_perf_model``).  Layers in FORWARD order;
    returns groups of layer indices in backward order."""
    L = len(sizes)
    p = list(sizes)
    tb = list(tb)
    Pm = world if world in perf_model.GbE_multi_p_ab_small else max(k for k in perf_model.GbE_multi_p_ab_small if k <= max(world, 2))
    topk = perf_model.topk_perf_model
    ag = lambda n: perf_model.allgather_perf_model(n, Pm, density)

    def schedule():
        ts = [topk(n) for n in p]
        tc = [ag(n) for n in p]
        taob, taos, taoc = [0.0] * L, [0.0] * L, [0.0] * L
        taos[L - 1] = taob[L - 1] + tb[L - 1]
        for l in range(L - 2, -1, -1):
            taob[l] = taos[l + 1] + ts[l + 1]
            taos[l] = taob[l] + tb[l]
        taoc[L - 1] = taos[L - 1] + ts[L - 1]
        for l in range(L - 2, -1, -1):
            taoc[l] = max(taoc[l + 1] + tc[l + 1], taos[l] + ts[l])
        return ts, taos, taoc

    ts, taos, taoc = schedule()
    groups, group = [], [L - 1] if L > 1 else []
    for l in range(L - 2, 0, -1):
        group.append(l)
        t_wait = tb[l - 1] + topk(p[l] + p[l - 1]) - topk(p[l]) - topk(p[l - 1]) - (taoc[l] - (taos[l] + ts[l]))
        t_save = ag(p[l]) + ag(p[l - 1]) - ag(p[l] + p[l - 1])
        if t_wait < t_save:
            tb[l - 1] += tb[l]; tb[l] = 0.0
            p[l - 1] += p[l]; p[l] = 0
            ts, taos, taoc = schedule()
        else:
            groups.append(group)
            group = []
    group.append(0)
    groups.append(group)
    # the walk above only closes groups between l and l-1 for l >= 2; de-duplicate and keep order
    seen, out = set(), []
    for g in groups:
        g2 = [i for i in g if i not in seen]
        seen.update(g2)
        if g2:
            out.append(g2)
    return out


class _DistributedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, named_parameters, compression=None, is_sparse=False, density=0.001,
                 seq_layernames=None, layerwise_times=None, norm_clip=None, threshold=0, fp16=False, mgwfbp=False,
                 asc=False, mgs=False, rdma=False, alpha=None, beta=None, verbose=True, momentum_correction=False,
                 profiling=False):
        super(self.__class__, self).__init__(params)
        # in-optimizer timers of the reference (wfbp/dopt.py:198-200,883-903; dead code there: the printer is a no-op):
        # host-side seconds per group for compression, collective launch and gradient write-back
        self._profiling = bool(profiling)
        self._compression_timers: Dict[str, list] = {}
        self._allreduce_timers: Dict[str, list] = {}
        self._update_times: Dict[str, list] = {}
        if not runtime.is_initialized():
            runtime.init()
        self._rank, self._world, self._device = runtime.rank(), runtime.size(), runtime.device()
        self._group = runtime.group()
        self._compression = compression or NoneCompressor()
        self._sparse = bool(is_sparse) and not isinstance(self._compression, NoneCompressor)
        self._density = density
        self._norm_clip = norm_clip
        self._gtopk = self._sparse and "gtopk" in getattr(self._compression, "name", "")
        self._mc = bool(momentum_correction) and self._sparse
        if self._mc and not isinstance(self, torch.optim.SGD):
            raise TypeError("momentum correction is defined for SGD with momentum (wfbp/dopt.py:906-953)")
        self._comm = None
        named = list(named_parameters)
        self._names = {p: n for n, p in named if p.requires_grad}
        self._params = [p for _, p in named if p.requires_grad]
        self._fwd_order = [n for n, p in named if p.requires_grad]
        self._cuda = self._device.type == "cuda"
        if self._cuda:
            self._stream = torch.cuda.Stream(device=self._device, priority=-1)
        self.alpha, self.beta = alpha, beta
        if mgs and self._sparse and layerwise_times is not None and seq_layernames is not None:
            by_name = {n: p for n, p in named}
            sizes = [by_name[n].numel() for n in seq_layernames]
            gidx = mgs_groups(sizes, layerwise_times, self._world, density)
            self._groups = [[seq_layernames[i] for i in g] for g in gidx]
        elif mgwfbp or asc:
            if layerwise_times is None or seq_layernames is None:
                raise ValueError("MG-WFBP needs seq_layernames and layerwise_times (utils.profiling.benchmark)")
            if self.alpha is None:
                table = perf_model.ALPHA_BETA_56GbIB if rdma else perf_model.ALPHA_BETA_10GbE
                self.alpha, self.beta = table.get(self._world, table[max(table)])
            by_name = {n: p for n, p in named}
            sizes = [by_name[n].numel() for n in seq_layernames]
            gidx = mgwfbp_groups(sizes, layerwise_times, self.alpha, self.beta, 2 if fp16 else 4, asc=asc)
            self._groups = [[seq_layernames[i] for i in g] for g in gidx]
        else:
            rev = self._fwd_order[::-1]
            by_name = {n: p for n, p in named}
            gidx = threshold_groups([by_name[n].numel() for n in rev], int(threshold))
            self._groups = [[rev[i] for i in g] for g in gidx]
        self._by_name = {n: p for n, p in named}
        self._group_of = {n: gi for gi, g in enumerate(self._groups) for n in g}
        self._buffers: Dict[int, torch.Tensor] = {}
        self._offsets: Dict[str, tuple] = {}
        for gi, g in enumerate(self._groups):
            off = 0
            for n in g:
                k = self._by_name[n].numel()
                self._offsets[n] = (off, off + k)
                off += k
            self._buffers[gi] = torch.zeros(off, device=self._device, dtype=self._by_name[g[0]].dtype)
        self._arrived = [0] * len(self._groups)
        self._launched: Dict[int, object] = {}
        self._hooks = []
        if self._world > 1:
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        if verbose and self._rank == 0:
            print("# of groups: ", len(self._groups), ", # of layers: ", len(self._params))

    # ---- alpha/beta measurement (wfbp/dopt.py:260-285) ----------------------------------------
    def benchmark_communication(self, num_iters: int = 20):
        from ...utils.profiling import CommunicationProfiler
        sync = (lambda: torch.cuda.synchronize()) if self._cuda else (lambda: None)
        prof = CommunicationProfiler(lambda t: dist.all_reduce(t, group=self._group), sync, device=self._device)
        sizes, times = prof.benchmark(num_iters)
        a, b = prof.fit_alpha_beta(sizes, times)
        ab = runtime.broadcast_object((a, b), src=0)
        self.alpha, self.beta = ab
        return ab

    # ---- backward hook ------------------------------------------------------------------------
    def _on_grad(self, p):
        n = self._names[p]
        gi = self._group_of[n]
        a, b = self._offsets[n]
        d_p = p.grad
        if self._mc:
            # momentum correction: sparsify the locally accumulated velocity, not the raw gradient
            st = self.state[p]
            buf = st.get("momentum_buffer")
            if buf is None:
                buf = st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            buf.mul_(self._momentum_of(p)).add_(d_p)
            d_p = buf
        self._buffers[gi][a:b].copy_(d_p.reshape(-1))
        self._arrived[gi] += 1
        if self._arrived[gi] == len(self._groups[gi]):
            self._launch(gi)

    def _momentum_of(self, p):
        for g in self.param_groups:
            if any(q is p for q in g["params"]):
                return g.get("momentum", 0.0)
        return 0.0

    def _tick(self, table, gi, t0):
        if self._profiling:
            import time
            if self._cuda:
                torch.cuda.synchronize(self._device)      # the reference's timers synchronise too (profiling mode only)
            table.setdefault("group-%d" % gi, []).append(time.perf_counter() - t0)

    def profiling_summary(self):
        """Mean seconds per group: {"compression": {...}, "allreduce": {...}, "update": {...}} (profiling=True)."""
        mean = lambda d: {k: sum(v) / len(v) for k, v in d.items() if v}
        return {"compression": mean(self._compression_timers), "allreduce": mean(self._allreduce_timers),
                "update": mean(self._update_times)}

    def _launch(self, gi):
        import time
        t_launch = time.perf_counter()
        buf = self._buffers[gi]
        ctx = torch.cuda.stream(self._stream) if self._cuda else None
        if self._cuda:
            self._stream.wait_stream(torch.cuda.current_stream(self._device))
            ctx.__enter__()
        try:
            if self._sparse:
                name = "group-%d" % gi
                _, idx, vals = self._compression.compress(buf, name, ratio=self._density)
                k = idx.numel()
                self._tick(self._compression_timers, gi, t_launch)
                if self._gtopk:
                    # global top-k of the SUM in log2(P) pairwise rounds; every rank ends with the same k entries
                    from ..comm import Comm
                    from .gtopk import gtopk_sparse_recursive_allreduce
                    if self._comm is None:
                        self._comm = Comm()
                    gv, gidx = gtopk_sparse_recursive_allreduce(self._comm, vals, idx, buf.numel(), k)
                    # locally selected values that did not survive the global cut return to the residual
                    lost = (~torch.isin(idx, gidx)).nonzero().view(-1)
                    keep = torch.ones(k, dtype=torch.bool, device=idx.device)
                    keep[lost] = False
                    self._compression.add_residuals(keep.nonzero().view(-1), name)
                    self._launched[gi] = ("gtopk", gv, gidx)
                    return
                all_vals = torch.empty(k * self._world, dtype=vals.dtype, device=buf.device)
                all_idx = torch.empty(k * self._world, dtype=idx.dtype, device=buf.device)
                dist.all_gather_into_tensor(all_vals, vals.contiguous(), group=self._group)
                dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=self._group)
                self._launched[gi] = (all_vals, all_idx)
            else:
                dist.all_reduce(buf, group=self._group)
                self._launched[gi] = True
            self._tick(self._allreduce_timers, gi, t_launch)
        finally:
            if self._cuda:
                ctx.__exit__(None, None, None)

    def synchronize(self):
        if self._world == 1:
            return
        for gi in range(len(self._groups)):          # groups whose parameters got no gradient
            if gi not in self._launched:
                self._launch(gi)
        if self._cuda:
            torch.cuda.current_stream(self._device).wait_stream(self._stream)
        import time
        for gi, g in enumerate(self._groups):
            t_up = time.perf_counter()
            buf = self._buffers[gi]
            res = self._launched.pop(gi)
            if self._sparse and res[0] == "gtopk":
                _, gv, gidx = res
                buf.zero_()
                buf.index_add_(0, gidx, gv.to(buf.dtype))
            elif self._sparse:
                all_vals, all_idx = res
                buf.zero_()
                buf.scatter_add_(0, all_idx, all_vals)
            buf.div_(self._world)
            for n in g:
                a, b = self._offsets[n]
                p = self._by_name[n]
                if p.grad is not None:
                    p.grad.copy_(buf[a:b].view_as(p.grad))
            if not self._sparse:
                buf.zero_()          # a parameter that misses its gradient next iteration must contribute zeros
            self._arrived[gi] = 0
            self._tick(self._update_times, gi, t_up)
        if self._norm_clip is not None:
            torch.nn.utils.clip_grad_norm_(self._params, self._norm_clip)

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.synchronize()
        if self._mc and self._world > 1:
            self._step_with_momentum_correction()
        else:
            super(self.__class__, self).step()
        return loss

    @torch.no_grad()
    def _step_with_momentum_correction(self):
        """wfbp/dopt.py:906-953: the communicated quantity already IS the (sparsified, averaged) velocity."""
        for group in self.param_groups:
            wd, lr = group["weight_decay"], group["lr"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                d_p = p.grad
                if wd != 0:
                    d_p = d_p.add(p, alpha=wd)
                p.add_(d_p, alpha=-lr)
                n = self._names.get(p)
                if n is None:
                    continue
                # momentum-factor masking: forget the velocity where it was just sent
                gi = self._group_of[n]
                zc = self._compression.zero_conditions.get("group-%d" % gi)
                buf = self.state[p].get("momentum_buffer")
                if zc is not None and buf is not None and self._density < 1:     # (reference: "and density < 1", :948)
                    a, b = self._offsets[n]
                    buf.view(-1).mul_(zc[a:b].to(buf.dtype))


def DistributedOptimizer(optimizer, named_parameters=None, model: Optional[nn.Module] = None, compression=None,
                         is_sparse=False, density=0.001, seq_layernames=None, layerwise_times=None, norm_clip=None,
                         threshold=0, writer=None, gradient_path=None, fp16=False, mgwfbp=False, asc=False, mgs=False,
                         rdma=False, multi_job_scheduling=False, alpha=None, beta=None, verbose=True,
                         momentum_correction=False, profiling=False, **ignored):
    """WFBP (``threshold=0``), threshold fusion, MG-WFBP (``mgwfbp=True``) or ASC (``asc=True``)."""
    if named_parameters is None:
        if model is None:
            raise ValueError("pass named_parameters or model")
        named_parameters = model.named_parameters()
    if isinstance(compression, str) or compression is None:
        compression = compressors[compression]()
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_DistributedOptimizer.__dict__))
    return cls(optimizer.param_groups, list(named_parameters), compression=compression, is_sparse=is_sparse,
               density=density, seq_layernames=seq_layernames, layerwise_times=layerwise_times, norm_clip=norm_clip,
               threshold=threshold, fp16=fp16, mgwfbp=mgwfbp, asc=asc, mgs=mgs, rdma=rdma, alpha=alpha, beta=beta,
               verbose=verbose, momentum_correction=momentum_correction, profiling=profiling)
