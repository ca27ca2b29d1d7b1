"""Sharded Adam / AdamW fused into the all-gather kernel (extension beyond the reference's SGD-only DeAR)."""
import os
import tempfile

import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model

CASES = [("adam", dict(lr=1e-2)), ("adam", dict(lr=1e-2, weight_decay=1e-2, betas=(0.8, 0.95))),
         ("adamw", dict(lr=1e-2, weight_decay=5e-2, eps=1e-6))]


def make_opt(kind, params, kw):
    return (torch.optim.AdamW if kind == "adamw" else torch.optim.Adam)(params, **kw)


def reference(kind, kw, steps, world, per_rank):
    model = make_model(); model.eval()
    opt = make_opt(kind, model.parameters(), kw)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    return [p.detach().clone() for p in model.parameters()]


def worker(rank, world, kind, kw, steps, per_rank, ckpt):
    import dear_pytorch_b200 as dear

    def train(model, opt, t0, t1):
        for t in range(t0, t1):
            x, y = data(t, world * per_rank)
            x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
            opt.zero_grad()
            nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
    model = make_model(); model.eval()
    opt = dear.DistributedOptimizer(make_opt(kind, model.parameters(), kw), model, threshold=0.002, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    half = steps // 2
    train(model, opt, 0, half)
    if ckpt:
        dear.save_checkpoint(ckpt, model, opt)
        opt.engine.close()
        model = make_model(seed=7); model.eval()
        opt = dear.DistributedOptimizer(make_opt(kind, model.parameters(), dict(lr=1.0)), model, threshold=None,
                                        num_nearby_layers=-1, verbose=False)
        dear.load_checkpoint(ckpt, model, opt)
    train(model, opt, half, steps)
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
@pytest.mark.parametrize("kind,kw", CASES)
def test_adam_matches_torch(backend, kind, kw):
    steps, world, per_rank = 6, 2, 4
    ref = reference(kind, kw, steps, world, per_rank)
    for params in run_ranks(worker, world=world, backend=backend, args=(kind, kw, steps, per_rank, None)):
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_adam_checkpoint_resume(backend):
    kind, kw = CASES[2]
    steps, world, per_rank = 6, 2, 4
    ref = reference(kind, kw, steps, world, per_rank)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "adam.pt")
        for params in run_ranks(worker, world=world, backend=backend, args=(kind, kw, steps, per_rank, path)):
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-6)
        sd = torch.load(path, weights_only=False)["optimizer"]
        stock = torch.optim.AdamW(make_model().parameters(), lr=1.0)
        stock.load_state_dict({"state": {k: {kk: vv for kk, vv in v.items() if kk != "master_param"} for k, v in sd["state"].items()},
                               "param_groups": sd["param_groups"]})       # loads into stock torch.optim.AdamW
        assert float(next(iter(stock.state.values()))["step"]) == 3.0


def gpu_worker(rank, world, kind, kw, steps, per_rank):
    import dear_pytorch_b200 as dear
    dev = dear.device()
    model = make_model().to(dev); model.eval()
    opt = dear.DistributedOptimizer(make_opt(kind, model.parameters(), kw), model, threshold=0.002, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank].to(dev), y[rank * per_rank:(rank + 1) * per_rank].to(dev)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    opt.synchronize()
    dear.communicator().check_status()
    return [p.detach().cpu() for p in model.parameters()]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_adam_on_gpu(world):
    kind, kw = CASES[2]
    ref = reference(kind, kw, 5, world, 4)
    for params in run_ranks(gpu_worker, world=world, backend="b200", args=(kind, kw, 5, 4),
                            extra_env={"DEAR_SPIN_TIMEOUT_S": "15"}, timeout=300):
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=2e-4, atol=5e-6)
