"""``Comm`` — one tensor-collective facade over the two data paths.

Same method family as the reference's pybind ``comm_core.Communicator``
(common/comm_core/src/comm_core.cpp:13-36): ``bcast, reduce, allReduce, allReduceRB, allReduceRSAG,
reduceScatter, allGather, sendrecv, multiBcast, synchronize, syncStream, getNumOfFreeStreams,
barrier`` — every asynchronous op returns a handle.

  * b200 / emu backends: the native ``_C.Communicator`` (our kernels over symmetric memory);
  * nccl / gloo backends: torch.distributed on a side stream (what the reference does over NCCL).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist

from .. import runtime


class Comm:
    def __init__(self, nstreams: int = 1):
        if not runtime.is_initialized():
            runtime.init()
        self.rank = runtime.rank()
        self.world = runtime.size()
        self.device = runtime.device()
        self.native = runtime.communicator()
        self.nstreams = max(1, nstreams)
        self._cuda = self.device.type == "cuda"
        if self.native is not None:
            # collective, like the reference's _extendComms (communicator.cpp:85-95): the shared native communicator
            # grows to the largest nstreams any Comm asked for; handles rotate over all of its slots
            self.native.extendStreams(self.nstreams)
        if self.native is None:
            self.group = runtime.group()
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.nstreams)] if self._cuda else []
            self._events = [torch.cuda.Event() for _ in range(self.nstreams)] if self._cuda else []
            self._cur = 0

    # ---- torch.distributed path -----------------------------------------------------------
    def _torch_run(self, fn: Callable) -> int:
        if self.world == 1:
            return 0
        if not self._cuda:
            fn()
            return 0
        i = self._cur
        self._cur = (self._cur + 1) % self.nstreams
        s = self._streams[i]
        s.wait_stream(torch.cuda.current_stream(self.device))     # the dependency the reference forgets
        with torch.cuda.stream(s):
            fn()
            self._events[i].record(s)
        return i

    # ---- collectives ----------------------------------------------------------------------
    def allReduce(self, tensor, scale: float = 1.0) -> int:
        if self.native is not None:
            return self.native.allReduce(tensor, scale)

        def f():
            dist.all_reduce(tensor, group=self.group)
            if scale != 1.0:
                tensor.mul_(scale)
        return self._torch_run(f)

    def allReduceRSAG(self, tensor, scale: float = 1.0) -> int:
        if self.native is not None:
            return self.native.allReduceRSAG(tensor, scale)
        return self.allReduce(tensor, scale)

    def allReduceRB(self, tensor, scale: float = 1.0) -> int:
        if self.native is not None:
            return self.native.allReduceRB(tensor, scale)

        def f():
            dist.reduce(tensor, dst=0, group=self.group)
            if scale != 1.0:
                tensor.mul_(scale)
            dist.broadcast(tensor, src=0, group=self.group)
        return self._torch_run(f)

    def reduce(self, tensor, root: int, scale: float = 1.0) -> int:
        if self.native is not None:
            return self.native.reduce(tensor, root, scale)

        def f():
            dist.reduce(tensor, dst=root, group=self.group)
            if scale != 1.0 and self.rank == root:
                tensor.mul_(scale)
        return self._torch_run(f)

    def bcast(self, tensor, root: int) -> int:
        if self.native is not None:
            return self.native.bcast(tensor, root)
        return self._torch_run(lambda: dist.broadcast(tensor, src=root, group=self.group))

    def reduceScatter(self, send, recv, scale: float = 1.0) -> int:
        if self.native is not None:
            return self.native.reduceScatter(send, recv, scale)

        def f():
            dist.reduce_scatter_tensor(recv, send, group=self.group)
            if scale != 1.0:
                recv.mul_(scale)
        if self.world == 1:
            recv.copy_(send).mul_(scale)
            return 0
        return self._torch_run(f)

    def allGather(self, send, recv) -> int:
        if self.native is not None:
            return self.native.allGather(send, recv)
        if self.world == 1:
            recv.copy_(send)
            return 0
        return self._torch_run(lambda: dist.all_gather_into_tensor(recv, send, group=self.group))

    def sendrecv(self, send, recv, peer: int) -> int:
        """Exchange with ``peer`` (every rank calls; peers must pair up)."""
        if self.native is not None:
            return self.native.sendrecv(send, recv, peer)

        def f():
            ops = [dist.P2POp(dist.isend, send, peer, group=self.group),
                   dist.P2POp(dist.irecv, recv, peer, group=self.group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return self._torch_run(f)

    def multiBcast(self, tensors: Sequence[torch.Tensor], outputs: Sequence[torch.Tensor], op: Callable) -> None:
        """Owner-computes-then-broadcasts for large tensors, everyone-computes for small ones
        (reference communicator.cpp:244-285; threshold 512x512 elements)."""
        owners = []
        k = 0
        for t, o in zip(tensors, outputs):
            if t.numel() >= 512 * 512 and self.world > 1:
                owner = k % self.world
                k += 1
                if owner == self.rank:
                    op(t, o)
                owners.append(owner)
            else:
                op(t, o)
                owners.append(-1)
        handles = [self.bcast(o, owner) for o, owner in zip(outputs, owners) if owner >= 0]
        for h in handles:
            self.waitStream(h)

    # ---- synchronisation --------------------------------------------------------------------
    def synchronize(self) -> None:
        if self.native is not None:
            self.native.synchronize()
        elif self._cuda:
            for s in self._streams:
                s.synchronize()

    def syncStream(self, handle: int) -> None:
        if self.native is not None:
            self.native.syncStream(handle)
        elif self._cuda:
            self._streams[handle].synchronize()

    def waitStream(self, handle: int) -> None:
        """Device-side dependency: the current stream waits for the op behind ``handle``."""
        if self.native is not None:
            self.native.waitStream(handle)
        elif self._cuda and self.world > 1:
            torch.cuda.current_stream(self.device).wait_event(self._events[handle])

    def getNumOfFreeStreams(self) -> int:
        if self.native is not None:
            return self.native.getNumOfFreeStreams()
        if not self._cuda:
            return self.nstreams
        return sum(1 for s in self._streams if s.query())

    def barrier(self) -> None:
        runtime.barrier()

    # ---- life cycle (reference Communicator::destroy / reload, communicator.cpp:43-83) ---------------
    def destroy(self) -> None:
        """Drain this communicator's streams.  The reference tears its NCCL communicators down here; the symmetric-memory
        runtime is shared by every ``Comm`` of the process and is released by ``dear.shutdown()``, so nothing is freed."""
        self.synchronize()
        self._destroyed = True

    def reload(self) -> None:
        """Counterpart of ``destroy``: make the communicator usable again (re-attaches to the runtime, which
        ``dear.shutdown()`` + ``dear.init()`` may have re-created in between)."""
        self.__init__(self.nstreams)
        self._destroyed = False
