"""Layer-wise backward profiler and communication (alpha-beta) profiler.

Reference: ``*/profiling.py`` — ``Profiling`` hooks every parameter gradient and forces a
``torch.cuda.synchronize()`` + ``time.time()`` per parameter (profiling.py:37-51);
``benchmark()`` = 5 warm-up + 50 timed iterations returning ``(seq_layernames, layerwise_times,
sizes)`` in forward order (profiling.py:98-129); ``CommunicationProfiler`` times a comm op over a
range of sizes (profiling.py:132-165) and the MG-WFBP optimizer fits alpha/beta by linear
regression (wfbp/dopt.py:260-285).

Here the per-layer timestamps are CUDA events recorded on the compute stream from
``register_post_accumulate_grad_hook`` — no device synchronisation inside the backward pass, so the
profile does not perturb what it measures; the events are resolved once per iteration.
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


class Profiling:
    def __init__(self, model: torch.nn.Module):
        self.model = model
        self._names = {p: n for n, p in model.named_parameters()}
        self._seq: List[str] = []               # backward order of the first iteration
        self._sizes: List[int] = []
        self._marks = {}                        # name -> [event or float] per iteration
        self._starts = []
        self._handles = []
        self._running = False
        self._cuda = next(model.parameters()).is_cuda
        for p in model.parameters():
            if p.requires_grad:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _now(self):
        if self._cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return time.perf_counter()

    def _hook(self, p):
        if not self._running:
            return
        name = self._names[p]
        if name not in self._marks:
            self._marks[name] = []
            self._seq.append(name)
            self._sizes.append(p.numel())
        self._marks[name].append(self._now())

    def start(self):
        """Call right before ``loss.backward()``."""
        self._running = True
        self._starts.append(self._now())

    def stop(self):
        self._running = False

    def reset(self):
        self._marks = {k: [] for k in self._marks}
        self._starts = []

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def get_backward_seq_keys(self):
        return list(self._seq)

    def get_backward_key_sizes(self):
        return list(self._sizes)

    def _delta(self, a, b) -> float:
        if self._cuda:
            return a.elapsed_time(b) / 1e3
        return b - a

    def get_layerwise_times(self) -> Tuple[List[float], float]:
        """Mean backward time attributed to each parameter (backward order) and mean total."""
        if self._cuda:
            torch.cuda.synchronize()
        n_iter = len(self._starts)
        per = np.zeros((n_iter, len(self._seq)))
        totals = np.zeros(n_iter)
        for it in range(n_iter):
            prev = self._starts[it]
            for j, name in enumerate(self._seq):
                if it >= len(self._marks[name]):
                    continue
                cur = self._marks[name][it]
                per[it, j] = max(self._delta(prev, cur), 0.0)
                prev = cur
            totals[it] = per[it].sum()
        return per.mean(0).tolist(), float(totals.mean())


def benchmark(model, fake_data, criterion, task: str = "imagenet", warmup: int = 5, iters: int = 50):
    """Return ``(seq_layernames, layerwise_times, sizes)`` in FORWARD order (reference profiling.py:98-129)."""
    if task == "bert":
        *inputs, target = fake_data
    else:
        inputs, target = [fake_data[0]], fake_data[1]
    p = Profiling(model)
    for i in range(warmup + iters):
        out = model(*inputs)
        loss = criterion(*out, *target) if isinstance(out, tuple) and isinstance(target, (tuple, list)) \
            else criterion(out, target)
        model.zero_grad(set_to_none=True)
        if i >= warmup:
            p.start()
        loss.backward()
        p.stop()
    times, _ = p.get_layerwise_times()
    keys, sizes = p.get_backward_seq_keys(), p.get_backward_key_sizes()
    p.close()
    return keys[::-1], times[::-1], sizes[::-1]


class CommunicationProfiler:
    """Time ``comm_op(tensor)`` followed by ``sync_op()`` over message sizes (elements)."""

    def __init__(self, comm_op: Callable, sync_op: Callable, sizes: Optional[Sequence[int]] = None, device=None):
        self.comm_op = comm_op
        self.sync_op = sync_op
        self.sizes = list(sizes) if sizes is not None else [2 ** 11 * i for i in range(1, 64)]   # 8 KB .. 504 KB fp32
        self.device = device

    def benchmark(self, num_iters: int = 100):
        elapsed = []
        for s in self.sizes:
            t = torch.rand(int(s), device=self.device)
            for _ in range(5):
                self.comm_op(t)
            self.sync_op()
            t0 = time.perf_counter()
            for _ in range(num_iters):
                self.comm_op(t)
            self.sync_op()
            elapsed.append((time.perf_counter() - t0) / num_iters)
        return list(self.sizes), elapsed

    @staticmethod
    def fit_alpha_beta(sizes_elems: Sequence[int], times: Sequence[float], bytes_per_elem: int = 4):
        """Least-squares fit  t = alpha + beta * bytes  (reference: sklearn LinearRegression)."""
        x = np.asarray(sizes_elems, dtype=np.float64) * bytes_per_elem
        y = np.asarray(times, dtype=np.float64)
        beta, alpha = np.polyfit(x, y, 1)
        return float(max(alpha, 0.0)), float(max(beta, 0.0))
