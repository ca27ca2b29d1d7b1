"""Host->device input pipeline.

``PinnedPrefetcher`` keeps ``depth`` batches in flight: while step *t* computes, batch *t+1* is
copied from pinned host memory into a **preallocated ring of device buffers** on a dedicated copy
stream, so the host->device transfer (38.5 MB per step for 64x3x224x224 fp32) is hidden behind
compute and the steady state performs no device allocation at all (an allocation per step on a side
stream makes the caching allocator fall back to cudaMalloc while the host runs ahead of the GPU).
The reference keeps one fixed batch on the device for the whole benchmark
(dear/imagenet_benchmark.py:97-103); this pipeline is what the end-to-end number of ``bench.py``
goes through.

Contract: a batch returned by ``next()`` stays valid until the work enqueued before the *next*
``next()`` call has consumed it (the usual "one batch per training step" loop).

``upload_delay_us``: a ring slot frees when the previous step's work retires, so every upload starts exactly at a
step boundary.  With the rotated training step (``TrainStep(overlap_update=True)``) the first thing a step runs is
Kernel B — the sharded update + all-gather, whose flag traffic uses system-scope release/acquire — and the PCIe DMA
landing at the same moment measurably stretches it (end-to-end minus device-resident step time: 0.01 ms with the
natural body, 0.18 ms with the rotated one at 1 GPU; profiles/bench_default_1gpu*.json).  A short spin on the COPY
stream in front of each upload moves the DMA into the forward pass, where the natural body already shows it is
free.  The copy still completes one and a half steps before its batch is consumed.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import torch


class SyntheticImages:
    """An endless stream of pinned host batches ``(images, labels)`` cycling over ``n_buffers``."""

    def __init__(self, batch_size: int, image_size: int = 224, num_classes: int = 1000, channels: int = 3,
                 n_buffers: int = 4, channels_last: bool = False, dtype=torch.float32, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.batches = []
        pin = torch.cuda.is_available()
        for _ in range(n_buffers):
            x = torch.randn(batch_size, channels, image_size, image_size, generator=g).to(dtype)
            if channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            y = torch.randint(0, num_classes, (batch_size,), generator=g)
            if pin:
                x, y = x.pin_memory(), y.pin_memory()
            self.batches.append((x, y))
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in self.batches[0])

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        i = 0
        while True:
            yield self.batches[i % len(self.batches)]
            i += 1


class PinnedPrefetcher:
    """Wrap an iterable of pinned host batches; yields device batches, ``depth`` copies ahead."""

    def __init__(self, host_batches: Iterable[Sequence[torch.Tensor]], device: torch.device, depth: int = 2,
                 upload_delay_us: float = 0.0):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth = max(1, depth)
        self.queue = []
        self.nslots = self.depth + 1
        self.ring = [None] * self.nslots          # device buffers, allocated once per slot
        self.free_ev = [None] * self.nslots       # compute-stream event: slot may be overwritten
        self._slot = 0
        self._last = None
        self._delay_cycles = 0
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device)
            if upload_delay_us > 0 and hasattr(torch.cuda, "_sleep"):
                try:
                    khz = getattr(torch.cuda.get_device_properties(self.device), "clock_rate", 0) or 1_900_000
                    with torch.cuda.stream(self.stream):
                        torch.cuda._sleep(1)                    # private torch API: probe it once, fall back to no delay
                    self._delay_cycles = int(upload_delay_us * khz / 1e3)
                except Exception:                               # pragma: no cover
                    self._delay_cycles = 0
        for _ in range(self.depth):
            self._enqueue()

    def _buffers_for(self, slot, host):
        bufs = self.ring[slot]
        if bufs is None or len(bufs) != len(host) or any(b.shape != h.shape or b.dtype != h.dtype or b.stride() != h.stride()
                                                         for b, h in zip(bufs, host)):
            bufs = tuple(torch.empty_strided(h.shape, h.stride(), dtype=h.dtype, device=self.device) for h in host)
            self.ring[slot] = bufs
        return bufs

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        if not self.cuda:
            self.queue.append((tuple(host), None, -1))
            return
        slot = self._slot
        self._slot = (slot + 1) % self.nslots
        bufs = self._buffers_for(slot, host)
        with torch.cuda.stream(self.stream):
            if self.free_ev[slot] is not None:
                self.stream.wait_event(self.free_ev[slot])      # the consumer is done with this slot
                if self._delay_cycles:
                    torch.cuda._sleep(self._delay_cycles)       # one spinning thread on the copy stream
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((bufs, ev, slot))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        if self.cuda and self._last is not None:
            # everything enqueued so far has consumed the previous batch: its slot is free after that
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self.free_ev[self._last] = done
        dev, ev, slot = self.queue.pop(0)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        self._last = slot if slot >= 0 else None
        self._enqueue()
        return dev
