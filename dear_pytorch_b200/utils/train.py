"""``TrainStep`` — one training iteration (zero_grad → forward → loss → backward → step) as a
callable, optionally captured into a single CUDA graph.

Why a graph: the reference's iteration issues ≈1.1 k tiny kernels and several host-blocking
stream synchronisations for ResNet-50.  With the fused kernels the optimizer
needs ~2 launches per bucket and no host synchronisation, and because the cross-GPU epochs live in
device memory (csrc/kernels.cu) the whole iteration — cuDNN/cuBLAS kernels, Kernel A per bucket on
the communication stream, Kernel B per bucket — replays as ONE graph launch.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch


class TrainStep:
    def __init__(self, model: torch.nn.Module, optimizer, loss_fn: Callable, autocast_dtype: Optional[torch.dtype] = None,
                 use_graph: bool = False, graph_warmup: int = 3):
        self.model = model
        self.opt = optimizer
        self.loss_fn = loss_fn
        self.autocast_dtype = autocast_dtype
        self.use_graph = use_graph
        self.graph_warmup = graph_warmup
        self._calls = 0
        self._graph = None
        self._static_in = None
        self._static_loss = None
        self._engine = getattr(optimizer, "_dear", None)
        self._debug = bool(os.environ.get("DEAR_GRAPH_DEBUG"))

    def _eager(self, *batch):
        *inputs, target = batch
        self.opt.zero_grad()
        if self.autocast_dtype is not None:
            with torch.autocast("cuda" if inputs[0].is_cuda else "cpu", dtype=self.autocast_dtype):
                out = self.model(*inputs)
        else:
            out = self.model(*inputs)
        loss = self.loss_fn(out, target)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def _log(self, msg):
        if self._debug:
            print("[TrainStep] " + msg, flush=True)

    def _capture(self, batch):
        eng = self._engine
        dev = batch[0].device
        self._static_in = tuple(torch.empty_like(t).copy_(t) for t in batch)
        # warm up on a side stream (the PyTorch whole-network capture recipe): allocator pools,
        # cuDNN autotuning and autograd streams all settle before the capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager(*self._static_in)
            if eng is not None:
                eng.synchronize(host=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        if eng is not None:
            eng.synchronize(host=True)
        torch.cuda.synchronize(dev)
        self._log("warm-up on side stream done; capturing")
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            loss = self._eager(*self._static_in)
            if eng is not None:
                eng.synchronize(host=False)     # join the communication stream back into the capture
            self._static_loss = loss
        self._log("capture done; first replay")
        # the capture only records; run the iteration for real
        self._graph.replay()
        return self._static_loss

    def __call__(self, *batch):
        self._calls += 1
        if not self.use_graph:
            return self._eager(*batch)
        if self._graph is None:
            if self._calls <= self.graph_warmup:
                return self._eager(*batch)
            return self._capture(batch)
        eng = self._engine
        if eng is not None and eng.hyper_changed():
            eng.refresh_hyper_outside_graph()
        for s, t in zip(self._static_in, batch):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self._graph.replay()
        return self._static_loss
