"""Stand-in for the reference's native ``comm_core`` extension (NCCL + MPI), which cannot be built
offline: it needs <mpi.h>/libmpi (absent from this image; `pip install` of common/comm_core fails at
`#include <mpi.h>`).  It issues the SAME NCCL collectives through torch.distributed,
one NCCL communicator + one side stream per ``Communicator`` like the original
(common/comm_core/src/communicator.cpp:43-66), with the same host-blocking ``synchronize()``
(communicator.cpp:103-110).  Process bootstrap comes from torchrun's env:// instead of mpirun.

This file is only used to run the reference's own Python (baseline/_ref/dear/dopt_rsag.py,
tensorfusion.py) unmodified as the comparison baseline.  None of the dear_pytorch_b200 engine is
involved.
"""
import datetime
import os

import torch
import torch.distributed as dist

_inited = False


def init():
    global _inited
    if _inited:
        return
    _inited = True
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=600))


def rank():
    return dist.get_rank() if dist.is_initialized() else int(os.environ.get("RANK", "0"))


def size():
    return dist.get_world_size() if dist.is_initialized() else int(os.environ.get("WORLD_SIZE", "1"))


def barriar():
    if dist.is_initialized():
        dist.barrier()


barrier = barriar


class Communicator(object):
    def __init__(self, nstreams=1):
        self._n = max(1, int(nstreams))
        self._multi = dist.is_initialized() and dist.get_world_size() > 1
        self._groups = [dist.new_group() if self._multi else None for _ in range(self._n)]
        self._streams = [torch.cuda.Stream() for _ in range(self._n)]
        self._cur = 0
        self._dirty = [False] * self._n

    def _next(self):
        i = self._cur
        self._cur = (self._cur + 1) % self._n
        return i

    def _run(self, fn, mark=True):
        i = self._next()
        if self._multi:
            # the original enqueues on its own stream without waiting for the compute stream
            with torch.cuda.stream(self._streams[i]):
                fn(self._groups[i])
        if mark:
            self._dirty[i] = True
        return i

    def destroy(self):
        pass

    def reload(self):
        pass

    def reduce(self, tensor, root):
        return self._run(lambda g: dist.reduce(tensor, dst=root, group=g))

    def bcast(self, tensor, root):
        return self._run(lambda g: dist.broadcast(tensor, src=root, group=g))

    def allReduce(self, tensor):
        self._run(lambda g: dist.all_reduce(tensor, group=g))

    def allReduceRB(self, tensor):
        def f(g):
            dist.reduce(tensor, dst=0, group=g)
            dist.broadcast(tensor, src=0, group=g)
        self._run(f)

    def allReduceRSAG(self, tensor):
        n = tensor.numel()
        P = size()
        if n < P:
            return self.allReduce(tensor)

        def f(g):
            pad = (P - n % P) % P
            buf = tensor.view(-1)
            if pad:
                buf = torch.zeros(n + pad, dtype=tensor.dtype, device=tensor.device)
                buf[:n].copy_(tensor.view(-1))
            shard = torch.empty((n + pad) // P, dtype=tensor.dtype, device=tensor.device)
            dist.reduce_scatter_tensor(shard, buf, group=g)
            dist.all_gather_into_tensor(buf, shard, group=g)
            if pad:
                tensor.view(-1).copy_(buf[:n])
        self._run(f)

    def reduceScatter(self, send, recv):
        self._run(lambda g: dist.reduce_scatter_tensor(recv, send, group=g), mark=False)

    def allGather(self, send, recv):
        self._run(lambda g: dist.all_gather_into_tensor(recv, send, group=g), mark=False)

    def sendrecv(self, send, recv, peer):
        def f(g):
            ops = [dist.P2POp(dist.isend, send, peer, group=g), dist.P2POp(dist.irecv, recv, peer, group=g)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self._run(f)

    def multiBcast(self, tensors, outputs, op):
        for t, o in zip(tensors, outputs):
            op(t, o)

    def synchronize(self):
        for s in self._streams:
            s.synchronize()
        self._dirty = [False] * self._n

    def syncStream(self, handle):
        if self._dirty[handle]:
            self._streams[handle].synchronize()
            self._dirty[handle] = False

    def getNumOfFreeStreams(self):
        return sum(1 for s in self._streams if s.query())

    def barrier(self):
        barriar()
