"""Host->device input pipeline.

``PinnedPrefetcher`` double-buffers batches: while step *t* computes, batch *t+1* is copied from
pinned host memory on a dedicated copy stream, so the host->device transfer (38.5 MB per step for
64x3x224x224 fp32) is hidden behind compute.  The reference keeps one fixed batch on the device
for the whole benchmark (dear/imagenet_benchmark.py:97-103); this pipeline is what the end-to-end
number of ``bench.py`` goes through.
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
    """Wrap an iterable of pinned host batches; yields device batches, one copy ahead."""

    def __init__(self, host_batches: Iterable[Sequence[torch.Tensor]], device: torch.device, depth: int = 2):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth = max(1, depth)
        self.queue = []
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device)
        for _ in range(self.depth):
            self._enqueue()

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        if not self.cuda:
            self.queue.append((tuple(host), None))
            return
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        dev, ev = self.queue.pop(0)
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev:
                t.record_stream(cur)      # allocated on the copy stream, consumed on the compute stream
        self._enqueue()
        return dev
