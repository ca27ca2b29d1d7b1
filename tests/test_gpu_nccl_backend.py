"""The engine on the ``nccl`` backend (parallel/backends.py: TorchBackend on CUDA streams) — the transport used when
a job spans several nodes (runtime.select_backend) and the reference's own transport.  world=1 runs on any GPU box;
the multi-rank cases need one physical GPU per rank (NCCL refuses two ranks on one device) and are skipped otherwise."""
import os

import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model, reference_run

pytestmark = pytest.mark.gpu


def nccl_worker(rank, world, case, steps, per_rank, method):
    import dear_pytorch_b200 as dear
    assert dear.backend() == "nccl"
    dev = dear.device()
    model = make_model().to(dev)
    model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, **case)
    if method == "dear":
        opt = dear.DistributedOptimizer(opt, model, threshold=0.001, verbose=False)
    else:
        from dear_pytorch_b200.parallel.baselines import wfbp
        opt = wfbp.DistributedOptimizer(opt, model.named_parameters(), threshold=2000)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank].to(dev), y[rank * per_rank:(rank + 1) * per_rank].to(dev)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    if hasattr(opt, "synchronize"):
        opt.synchronize()
    torch.cuda.synchronize()
    return [p.detach().float().cpu() for p in model.parameters()]


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [1, 2, 4])
def test_dear_engine_on_nccl_backend(world):
    if world > 1 and _gpus() < world:
        pytest.skip("needs %d physical GPUs" % world)
    case = dict(momentum=0.9, weight_decay=1e-3)
    ref = reference_run(case, 4, world, 4)
    for params in run_ranks(nccl_worker, world=world, backend="nccl", args=(case, 4, 4, "dear"), timeout=300):
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


@pytest.mark.multigpu
def test_wfbp_baseline_on_nccl():
    if _gpus() < 2:
        pytest.skip("needs 2 physical GPUs")
    case = dict(momentum=0.9)
    ref = reference_run(case, 3, 2, 4)
    for params in run_ranks(nccl_worker, world=2, backend="nccl", args=(case, 3, 4, "wfbp"), timeout=300):
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
