"""Tensor-fusion helper library (name-keyed fusion buffers and merged collectives).

Counterpart of the reference's dear/tensorfusion.py:
  ``CollectiveOp``            :204-209
  ``TensorGroup``             :14-200   name -> (group, slot) map, lazily allocated flat buffers,
                                        push (copy-in, "group full" detection), pull, regrouping,
                                        wait-in-buffer time accounting
  ``MergedCommCollective``    :211-349  merged reduce / bcast / all-reduce with optional
  ``MergedCommReduce``        :352-453  ``symmetric`` upper-triangle packing
  ``CommReduceScatter``       :455-487  RS / AG dispatch used by the DeAR optimizer
The DeAR engine itself (parallel/optimizer.py) does not go through these classes — its buckets are
symmetric-memory views driven by the fused kernels — they serve the reduce/broadcast variant, the
baselines and user code written against the reference's helper API.
"""
from __future__ import annotations

import time
from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .comm import Comm


class CollectiveOp(Enum):
    REDUCE = 0
    BCAST = 1
    ALL_REDUCE = 2
    REDUCE_SCATTER = 3
    ALL_GATHER = 4


class TensorGroup:
    """Groups named tensors into flat fusion buffers."""

    def __init__(self, tensor_names: Sequence[str], single_layer: bool = False, tensors: Optional[Dict] = None,
                 num_nearby_layers: int = 4, threshold_mb: Optional[float] = None, sizes: Optional[Dict[str, int]] = None):
        self._names = list(tensor_names)
        self._sizes = dict(sizes or {})
        if tensors:
            self._sizes.update({n: t.numel() for n, t in tensors.items()})
        if single_layer:
            groups = [[n] for n in self._names]
        elif threshold_mb is not None and self._sizes:
            groups = self._group_by_threshold(threshold_mb)
        else:
            groups = self._group_by_count(num_nearby_layers)
        self._set_groups(groups)
        self._buffers: Dict[int, torch.Tensor] = {}
        self._in_times: Dict[str, float] = {}
        self.wait_times: Dict[str, float] = {n: 0.0 for n in self._names}

    # ---- grouping ---------------------------------------------------------------------------
    def _group_by_count(self, k: int) -> List[List[str]]:
        if k < 0:
            return [list(self._names)]
        return [self._names[i:i + k] for i in range(0, len(self._names), max(1, k))]

    def _group_by_threshold(self, threshold_mb: float) -> List[List[str]]:
        groups, cur, tot = [], [], 0.0
        for n in self._names:
            sz = self._sizes.get(n, 0) * 4 / 1024 / 1024
            if tot == 0 or tot + sz < threshold_mb:
                cur.append(n)
                tot += sz
            else:
                groups.append(cur)
                cur, tot = [n], sz
        if cur:
            groups.append(cur)
        return groups

    def _set_groups(self, groups: List[List[str]]):
        self._groups = groups
        self._where: Dict[str, Tuple[int, int]] = {}
        for gi, g in enumerate(groups):
            for si, n in enumerate(g):
                self._where[n] = (gi, si)
        self._flags = [[0] * len(g) for g in groups]
        self._offsets: Dict[str, Tuple[int, int]] = {}
        self._buffers = {}

    def regroup_by_flags(self, flags: Sequence[int]):
        """``flags[i] == 1`` starts a new group at tensor i (reference :157-175)."""
        groups, cur = [], []
        for n, f in zip(self._names, flags):
            if f and cur:
                groups.append(cur)
                cur = []
            cur.append(n)
        if cur:
            groups.append(cur)
        self._set_groups(groups)

    def regroup(self, groups: List[List[str]]):
        self._set_groups([list(g) for g in groups])

    @property
    def groups(self) -> List[List[str]]:
        return self._groups

    def is_merged(self) -> bool:
        return any(len(g) > 1 for g in self._groups)

    def get_group_index_by_name(self, name: str) -> Tuple[int, int]:
        return self._where[name]

    # ---- buffers ------------------------------------------------------------------------------
    def _ensure_buffer(self, gi: int, like: torch.Tensor) -> torch.Tensor:
        if gi not in self._buffers:
            off = 0
            for n in self._groups[gi]:
                sz = self._sizes[n]
                self._offsets[n] = (off, off + sz)
                off += sz
            self._buffers[gi] = like.new_zeros(off)
        return self._buffers[gi]

    def push_tensor(self, name: str, tensor: torch.Tensor):
        """Copy ``tensor`` into its slot; returns ``(group_name, flat_buffer)`` when the group is
        complete, else ``(None, None)``."""
        gi, si = self._where[name]
        self._sizes.setdefault(name, tensor.numel())
        missing = [n for n in self._groups[gi] if n not in self._sizes]
        if missing:
            raise KeyError("sizes of %s must be known before the first push" % missing)
        buf = self._ensure_buffer(gi, tensor)
        a, b = self._offsets[name]
        buf[a:b].copy_(tensor.reshape(-1))
        self._flags[gi][si] = 1
        self._in_times[name] = time.perf_counter()
        if not all(self._flags[gi]):
            return None, None
        now = time.perf_counter()
        for n in self._groups[gi]:
            self.wait_times[n] = 0.1 * self.wait_times[n] + 0.9 * (now - self._in_times[n]) * 1e3
        return "group-%d" % gi, buf

    def pull_alltensors(self) -> Dict[str, torch.Tensor]:
        out = {}
        for gi, g in enumerate(self._groups):
            if gi not in self._buffers:
                continue
            for n in g:
                a, b = self._offsets[n]
                out[n] = self._buffers[gi][a:b]
        return out

    def buffer(self, gi: int) -> torch.Tensor:
        return self._buffers[gi]

    def clear_flags(self):
        self._flags = [[0] * len(g) for g in self._groups]


def _triu_pack(t: torch.Tensor) -> torch.Tensor:
    idx = torch.triu_indices(t.shape[0], t.shape[1], device=t.device)
    return t[idx[0], idx[1]]


def _triu_unpack(flat: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    idx = torch.triu_indices(out.shape[0], out.shape[1], device=out.device)
    out[idx[0], idx[1]] = flat
    out[idx[1], idx[0]] = flat
    return out


class MergedCommCollective:
    """Merged (fused) reduce / broadcast / all-reduce over named tensors."""

    def __init__(self, tensor_names: Optional[Sequence[str]] = None, merge: bool = True, single_layer: bool = False,
                 symmetric: bool = False, op: CollectiveOp = CollectiveOp.ALL_REDUCE, num_nearby_layers: int = 4,
                 nstreams: int = 1, comm: Optional[Comm] = None):
        self.op = op
        self.merge = merge
        self.symmetric = symmetric
        self.single_layer = single_layer
        self.num_nearby_layers = num_nearby_layers
        self.merged_comm = comm or Comm(nstreams)
        self._group: Optional[TensorGroup] = None
        self._name_tensors: Dict[str, torch.Tensor] = {}
        self.handles: List[Tuple[int, str]] = []
        if tensor_names is not None:
            self.init_tensor_group(tensor_names)

    def init_tensor_group(self, tensor_names: Sequence[str], sizes: Optional[Dict[str, int]] = None):
        if self.merge:
            self._group = TensorGroup(tensor_names, single_layer=self.single_layer,
                                      num_nearby_layers=self.num_nearby_layers, sizes=sizes)

    def _issue(self, tensor: torch.Tensor, root: int) -> int:
        scale = 1.0
        if self.op == CollectiveOp.REDUCE:
            return self.merged_comm.reduce(tensor, root, scale)
        if self.op == CollectiveOp.BCAST:
            return self.merged_comm.bcast(tensor, root)
        if self.op == CollectiveOp.ALL_REDUCE:
            return self.merged_comm.allReduce(tensor, scale)
        raise TypeError("unsupported op %s" % self.op)

    def collective_async_(self, name: str, tensor: torch.Tensor, root: int = 0):
        """Returns a handle if a collective was launched, else ``None`` (tensor only buffered)."""
        self._name_tensors[name] = tensor
        payload = _triu_pack(tensor) if (self.symmetric and tensor.dim() == 2 and tensor.shape[0] == tensor.shape[1]) \
            else tensor.reshape(-1)
        if self.merge and self._group is not None:
            self._group._sizes.setdefault(name, payload.numel())
            gname, buf = self._group.push_tensor(name, payload)
            if buf is None:
                return None
            h = self._issue(buf, root)
            self.handles.append((h, gname))
            return h
        if payload.data_ptr() != tensor.data_ptr():
            self._packed = getattr(self, "_packed", {})
            self._packed[name] = payload
        h = self._issue(payload if payload.is_contiguous() else payload.contiguous(), root)
        self.handles.append((h, name))
        return h

    def synchronize(self) -> Dict[str, torch.Tensor]:
        """Host-blocking, like the reference (:329-349): waits, scatters merged results back."""
        self.merged_comm.synchronize()
        if self.merge and self._group is not None:
            for n, flat in self._group.pull_alltensors().items():
                t = self._name_tensors.get(n)
                if t is None:
                    continue
                if self.symmetric and t.dim() == 2 and t.shape[0] == t.shape[1] and flat.numel() != t.numel():
                    _triu_unpack(flat, t)
                else:
                    t.reshape(-1).copy_(flat)
            self._group.clear_flags()
        else:
            for n, flat in getattr(self, "_packed", {}).items():
                _triu_unpack(flat, self._name_tensors[n])
            self._packed = {}
        out = dict(self._name_tensors)
        self._name_tensors = {}
        self.handles = []
        return out


class MergedCommReduce(MergedCommCollective):
    """Reduce-to-root flavour used by the reduce/broadcast variant (reference :352-453)."""

    def __init__(self, tensor_names=None, merge=True, single_layer=False, symmetric=False,
                 op: CollectiveOp = CollectiveOp.REDUCE, **kw):
        super().__init__(tensor_names, merge=merge, single_layer=single_layer, symmetric=symmetric, op=op, **kw)

    def reduce_async_(self, name: str, tensor: torch.Tensor, root_rank: int = 0):
        return self.collective_async_(name, tensor, root_rank)


class CommReduceScatter:
    """RS / AG dispatcher with the reference's interface (dear/tensorfusion.py:455-487)."""

    def __init__(self, tensor_names=None, op: CollectiveOp = CollectiveOp.REDUCE_SCATTER, comm: Optional[Comm] = None):
        self.op = op
        self.merged_comm = comm or Comm(1)
        self._name_tensors = {}
        self.handles = []

    def init_tensor_group(self, tensor_names, num_nearby_layers=4):
        pass

    def collective_async_(self, name, pad_tensor, shard_tensor):
        self._name_tensors[name] = (pad_tensor, shard_tensor)
        if self.op == CollectiveOp.REDUCE_SCATTER:
            h = self.merged_comm.reduceScatter(pad_tensor, shard_tensor)
        elif self.op == CollectiveOp.ALL_GATHER:
            h = self.merged_comm.allGather(shard_tensor, pad_tensor)
        else:
            raise TypeError
        self.handles.append((h, shard_tensor, pad_tensor))
        return h

    def synchronize(self):
        self.merged_comm.synchronize()
        self._name_tensors.clear()
        self.handles.clear()
