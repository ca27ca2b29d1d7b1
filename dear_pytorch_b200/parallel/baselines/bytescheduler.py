"""ByteScheduler-style communication scheduling on torch.distributed (ByteScheduler is not installable here).

Reference driver: bytescheduler/imagenet_benchmark.py:73-82 (``bsc.ScheduledOptimizer(model, hvd_optimizer, steps)``)
launched with ``BYTESCHEDULER_PARTITION=4000000  BYTESCHEDULER_CREDIT=16000000  BYTESCHEDULER_CREDIT_TUNING=0``
(bytescheduler/horovod_mpi_cj.sh:24-27).  What that library does, and what is rebuilt here:

  * **tensor partitioning** — every gradient is cut into chunks of at most ``partition`` elements; each chunk is an
    independent all-reduce, so a large tensor can be pre-empted between chunks;
  * **priority scheduling** — chunks are served in FORWARD order (the parameter the next forward pass needs first has
    the highest priority), not in the order back-propagation produced them;
  * **credit** — at most ``credit`` elements are committed to the (FIFO) communication stream at any time; everything
    else waits in the priority queue, which is what lets a late high-priority gradient overtake earlier ones;
  * **cross-iteration overlap** — ``step()`` does not wait: a module's forward pre-hook waits for *its own*
    parameters' chunks and applies their update (per-parameter SGD / Adam, as ByteScheduler's ``_sgd`` / ``_adam``),
    so low-priority communication overlaps the next forward pass.

NCCL needs the same collective order on every rank, which Horovod's coordinator provides for the original.  Here the
schedule is a deterministic function of the gradient arrival order (identical on every rank): credit is only ever
returned by waiting for the OLDEST in-flight chunk, never by polling completion.  During back-propagation the hook
launches while credit lasts and waits for at most one chunk per gradient; after back-propagation a scheduler thread
drains the queue in priority order while the main thread already runs the next forward pass.
"""
from __future__ import annotations

import collections
import heapq
import os
import threading
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import runtime


def partition_sizes(numel: int, partition: int) -> List[int]:
    """Chunk sizes of one tensor: ``partition`` elements each, the remainder last; ``partition <= 0`` = no split."""
    if partition <= 0 or numel <= partition:
        return [numel]
    full, rest = divmod(numel, partition)
    return [partition] * full + ([rest] if rest else [])


def _flat_dense(t: torch.Tensor) -> torch.Tensor:
    if t.is_contiguous():
        return t.view(-1)
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())   # dense, permuted (channels-last)


class _Chunk:
    __slots__ = ("prio", "idx", "tensor", "param", "done")

    def __init__(self, prio, idx, tensor, param):
        self.prio, self.idx, self.tensor, self.param, self.done = prio, idx, tensor, param, None

    def __lt__(self, other):
        return (self.prio, self.idx) < (other.prio, other.idx)


class _ByteSchedulerOptimizer(torch.optim.Optimizer):
    def __init__(self, params, model, partition=None, credit=None, verbose=True):
        super(self.__class__, self).__init__(params)
        if not runtime.is_initialized():
            runtime.init()
        self._rank, self._world, self._device = runtime.rank(), runtime.size(), runtime.device()
        # the scheduler thread issues collectives while the main thread runs user code: they travel on a communicator
        # of their own (as ByteScheduler's / Horovod's do), so a user collective on the default group — a metric
        # all-reduce, a barrier — can never interleave differently on two ranks
        self._pg = runtime.group()
        if self._world > 1 and self._pg is None:
            self._pg = dist.new_group(ranks=list(range(self._world)))
        self.partition = int(partition if partition is not None else os.environ.get("BYTESCHEDULER_PARTITION", 4000000))
        self.credit = int(credit if credit is not None else os.environ.get("BYTESCHEDULER_CREDIT", 16000000))
        self._cuda = self._device.type == "cuda"
        self._stream = torch.cuda.Stream(device=self._device, priority=-1) if self._cuda else None
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._prio = {p: i for i, p in enumerate(self._params)}          # forward order: 0 = needed first
        self._group_of = {p: g for g in self.param_groups for p in g["params"]}
        self._gidx = {p: i for i, g in enumerate(self.param_groups) for p in g["params"]}
        self._hyper_of = {}                 # deferred parameter -> hyper-parameters as of the step() that deferred it
        self._lazy = isinstance(self, (torch.optim.SGD, torch.optim.Adam, torch.optim.AdamW))
        self._heap: List[_Chunk] = []
        self._inflight = collections.deque()                               # chunks committed to the comm stream
        self._inflight_elems = 0
        self._chunks: Dict[nn.Parameter, List[_Chunk]] = {}               # gradient being synchronised, per parameter
        self._launched_all: Dict[nn.Parameter, threading.Event] = {}
        self._left: Dict[nn.Parameter, int] = {}
        self._thread: Optional[threading.Thread] = None
        self._thread_error = None
        self._lock = threading.Lock()
        self.launch_log: List[tuple] = []                                  # (priority, chunk index, numel): tests / traces
        self._hooks = []
        if self._world > 1:
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            for m in model.modules():
                mine = [p for p in m.parameters(recurse=False) if p.requires_grad]
                if mine:
                    self._hooks.append(m.register_forward_pre_hook(self._make_pre_hook(mine)))
        if verbose and self._rank == 0:
            print("[bytescheduler-like] partition %d elements, credit %d elements, %d tensors" % (
                self.partition, self.credit, len(self._params)))

    # ---- launching / credit ---------------------------------------------------------------------
    def _launch(self, c: _Chunk):
        if self._cuda:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(c.tensor, group=self._pg)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            c.tensor.record_stream(self._stream)
            c.done = ev
        else:
            c.done = dist.all_reduce(c.tensor, group=self._pg, async_op=True)
        self._inflight.append(c)
        self._inflight_elems += c.tensor.numel()
        self.launch_log.append((c.prio, c.idx, c.tensor.numel()))
        with self._lock:
            self._left[c.param] -= 1
            if self._left[c.param] == 0:
                self._launched_all[c.param].set()

    def _retire_oldest(self):
        c = self._inflight.popleft()
        if self._cuda:
            c.done.synchronize()
        else:
            c.done.wait()
        self._inflight_elems -= c.tensor.numel()

    def _fits(self, c: _Chunk) -> bool:
        return not self._inflight or self._inflight_elems + c.tensor.numel() <= self.credit

    def _pump(self, may_block_once: bool):
        """Launch queued chunks in priority order while credit lasts; optionally free credit by waiting for the
        oldest in-flight chunk ONCE (backward hooks), so the host never stalls for long inside autograd."""
        while self._heap:
            if self._fits(self._heap[0]):
                self._launch(heapq.heappop(self._heap))
            elif may_block_once:
                self._retire_oldest()
                may_block_once = False
            else:
                break

    def _drain(self):
        try:
            if self._cuda:
                torch.cuda.set_device(self._device)
            while self._heap:
                while not self._fits(self._heap[0]):
                    self._retire_oldest()
                self._launch(heapq.heappop(self._heap))
        except BaseException as e:          # surfaced by the next join()
            self._thread_error = e
            for ev in self._launched_all.values():
                ev.set()

    def _join(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
            if self._thread_error is not None:
                err, self._thread_error = self._thread_error, None
                raise err

    # ---- hooks ----------------------------------------------------------------------------------
    def _on_grad(self, p):
        self._join()                                   # the previous step's scheduler thread owns the queue until done
        if p in self._chunks:
            # its module's forward pre-hook settles a parameter before the parameter can produce a new gradient; a
            # parameter that is used OUTSIDE the module that owns it (functional call, tied weight) escapes that
            raise RuntimeError("ByteSchedulerOptimizer: a new gradient arrived for a parameter whose previous gradient "
                               "is still being synchronised; parameters must be used by the module that owns them")
        flat = _flat_dense(p.grad)
        sizes = partition_sizes(flat.numel(), self.partition)
        if self._cuda:
            self._stream.wait_stream(torch.cuda.current_stream(self._device))     # the gradient is being produced there
        chunks, off = [], 0
        for i, n in enumerate(sizes):
            chunks.append(_Chunk(self._prio[p], i, flat[off:off + n], p))
            off += n
        self._chunks[p] = chunks
        self._left[p] = len(chunks)
        self._launched_all[p] = threading.Event()
        for c in chunks:
            heapq.heappush(self._heap, c)
        self._pump(may_block_once=True)

    def _make_pre_hook(self, params):
        def hook(module, inputs):
            for p in params:
                if p in self._chunks:
                    self._finish(p)
        return hook

    def _finish(self, p):
        """Wait for p's chunks (stream-wise on GPU), average, and — lazy mode — apply p's update."""
        if not self._launched_all[p].is_set() and self._thread is None:
            # a forward pass between backward() and step() (no scheduler thread yet): schedule what is queued now,
            # in the same deterministic priority / credit order the thread would use
            self._drain()
            if self._thread_error is not None:
                err, self._thread_error = self._thread_error, None
                raise err
        self._launched_all[p].wait()
        if self._thread_error is not None:
            self._join()
        for c in self._chunks.pop(p):
            if self._cuda:
                torch.cuda.current_stream(self._device).wait_event(c.done)
            else:
                c.done.wait()
        del self._launched_all[p], self._left[p]
        p.grad.div_(self._world)
        if self._lazy and p in self._deferred:
            self._deferred.discard(p)
            self._update_one(p)
            p.grad = None

    # ---- per-parameter updates (ByteScheduler's _sgd / _adam) ---------------------------------------
    @torch.no_grad()
    def _update_one(self, p):
        # a deferred update belongs to the step() that scheduled it: an LR scheduler may have moved on since
        snap = self._hyper_of.pop(p, None)
        g = snap[self._gidx[p]] if snap is not None else self._group_of[p]
        d = p.grad
        st = self.state[p]
        if isinstance(self, torch.optim.SGD):
            wd, mom, damp, nest = g["weight_decay"], g["momentum"], g["dampening"], g["nesterov"]
            if wd != 0:
                d = d.add(p, alpha=wd)
            if mom != 0:
                buf = st.get("momentum_buffer")
                if buf is None:
                    buf = st["momentum_buffer"] = torch.clone(d).detach()
                else:
                    buf.mul_(mom).add_(d, alpha=1 - damp)
                d = d.add(buf, alpha=mom) if nest else buf
            p.add_(d, alpha=-g["lr"])
            return
        b1, b2 = g["betas"]
        if "step" not in st:
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st["step"] += 1
        t = float(st["step"])
        if isinstance(self, torch.optim.AdamW):
            p.mul_(1 - g["lr"] * g["weight_decay"])
        elif g["weight_decay"] != 0:
            d = d.add(p, alpha=g["weight_decay"])
        st["exp_avg"].mul_(b1).add_(d, alpha=1 - b1)
        st["exp_avg_sq"].mul_(b2).addcmul_(d, d, value=1 - b2)
        denom = (st["exp_avg_sq"].sqrt() / (1 - b2 ** t) ** 0.5).add_(g["eps"])
        p.addcdiv_(st["exp_avg"], denom, value=-g["lr"] / (1 - b1 ** t))

    # ---- optimizer API ----------------------------------------------------------------------------
    _deferred: set = set()

    def synchronize(self):
        """Wait for every outstanding chunk and apply every deferred update."""
        if self._world == 1:
            return
        self._join()
        self._pump(may_block_once=False)
        while self._heap:
            self._retire_oldest()
            self._pump(may_block_once=False)
        for p in list(self._chunks):
            self._finish(p)
        while self._inflight:
            self._retire_oldest()

    def zero_grad(self, set_to_none: bool = True):
        """Gradients still being synchronised are left alone (they are released when their update is applied)."""
        for p in self._params:
            if p in self._chunks or p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.zero_()

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._world == 1:
            super(self.__class__, self).step()
            return loss
        if not self._lazy:
            self.synchronize()
            super(self.__class__, self).step()
            return loss
        # lazy: every parameter with a gradient in flight is updated when its module next runs (or at synchronize())
        del self.launch_log[:-4096]
        snap = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        self._deferred = set(self._chunks)
        for p in self._deferred:            # (a module that does not run next iteration keeps ITS step's values)
            self._hyper_of.setdefault(p, snap)
        for p in self._params:                      # a gradient that never went through the hook (no peers to wait for)
            if p.grad is not None and p not in self._chunks:
                self._update_one(p)
                p.grad = None
        self._thread = threading.Thread(target=self._drain, name="bytescheduler-drain", daemon=True)
        self._thread.start()
        return loss


def ByteSchedulerOptimizer(optimizer, model: nn.Module, partition: Optional[int] = None, credit: Optional[int] = None,
                           verbose: bool = True, **ignored):
    """``bsc.ScheduledOptimizer`` look-alike: partitioned, priority-scheduled, credit-limited all-reduces with the
    update of every parameter deferred to its module's next forward (see the module docstring)."""
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_ByteSchedulerOptimizer.__dict__))
    return cls(optimizer.param_groups, model, partition=partition, credit=credit, verbose=verbose)
