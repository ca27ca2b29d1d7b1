"""gTop-k sparse all-reduce (Shi et al., ICDCS 2019): log2(P) rounds of pairwise exchange of
(values, indices) followed by a local re-selection of the k largest magnitudes.

Reference: wfbp/dopt.py:50-106 (``gtopk_sparse_recursive_allreduce`` on comm_core.sendrecv, which
needs the missing ``tcmm`` extension).  Here the exchange is ``Comm.sendrecv`` (our peer-copy kernel
on the b200 backend, NCCL/gloo point-to-point otherwise) and the merge is plain torch.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

from ..comm import Comm


@torch.no_grad()
def gtopk_sparse_recursive_allreduce(comm: Comm, values: torch.Tensor, indexes: torch.Tensor, numel: int, k: int
                                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks end with the same global top-``k`` ``(values, indexes)`` of the SUM of the sparse
    inputs.  ``world`` must be a power of two; every rank passes exactly ``k`` entries."""
    world, rank = comm.world, comm.rank
    if world & (world - 1):
        raise ValueError("gTop-k needs a power-of-two world size")
    vals = values.clone().float()
    idx = indexes.clone().long()
    rounds = int(math.log2(world)) if world > 1 else 0
    for r in range(rounds):
        peer = rank ^ (1 << r)
        send = torch.cat([vals, idx.to(vals.dtype)])        # indices < 2**24 are exact in fp32
        if numel >= 2 ** 24:
            send = torch.cat([vals.double(), idx.double()])
        recv = torch.empty_like(send)
        comm.waitStream(comm.sendrecv(send, recv, peer))
        comm.syncStream(0) if comm.native is None and comm.device.type == "cuda" else None
        pv, pi = recv[:k].to(vals.dtype), recv[k:].long()
        dense = torch.zeros(numel, dtype=vals.dtype, device=vals.device)
        dense.index_add_(0, idx, vals)
        dense.index_add_(0, pi, pv)
        cand = torch.unique(torch.cat([idx, pi]))
        cv = dense[cand]
        if cand.numel() > k:
            top = torch.topk(cv.abs(), k, sorted=False)[1]
            cand, cv = cand[top], cv[top]
        elif cand.numel() < k:                                # pad so that message sizes stay fixed
            pad = k - cand.numel()
            cand = torch.cat([cand, cand.new_zeros(pad)])
            cv = torch.cat([cv, cv.new_zeros(pad)])
        order = torch.argsort(cand)
        idx, vals = cand[order], cv[order]
    return vals, idx
