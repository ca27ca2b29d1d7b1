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


class _Branchy(nn.Module):
    """``side`` only takes part in odd steps: torch.optim.Adam then counts ITS steps separately (and leaves it alone
    otherwise)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.a, self.side, self.b = nn.Linear(10, 16), nn.Linear(16, 16), nn.Linear(16, 4)

    def forward(self, x, use_side):
        h = torch.tanh(self.a(x))
        if use_side:
            h = h + torch.relu(self.side(h))
        return self.b(h)


def _branchy_data(t, n):
    g = torch.Generator().manual_seed(900 + t)
    return torch.randn(n, 10, generator=g), torch.randint(0, 4, (n,), generator=g)


def _branchy_reference(kind, kw, steps, n):
    m = _Branchy()
    opt = make_opt(kind, m.parameters(), kw)
    for t in range(steps):
        x, y = _branchy_data(t, n)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x, t % 3 == 2), y).backward()
        opt.step()
    return [p.detach().clone() for p in m.parameters()], {i: int(st["step"]) for i, st in opt.state_dict()["state"].items()}


def _branchy_worker(rank, world, kind, kw, steps, per, reload_at):
    import dear_pytorch_b200 as dear
    m = _Branchy()
    opt = dear.DistributedOptimizer(make_opt(kind, m.parameters(), kw), m, threshold=0.0005, verbose=False)
    dear.broadcast_parameters(m.state_dict(), 0)
    for t in range(steps):
        x, y = _branchy_data(t, world * per)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x[rank * per:(rank + 1) * per], t % 3 == 2), y[rank * per:(rank + 1) * per]).backward()
        opt.step()
        if t == reload_at:                           # the per-parameter counts survive a state-dict round trip
            opt.load_state_dict(opt.state_dict())
    opt.synchronize()
    sd = opt.state_dict()
    return [p.detach().clone() for p in m.parameters()], {i: int(float(e["step"])) for i, e in sd["state"].items() if "step" in e}


@pytest.mark.parametrize("backend", ["gloo", "emu"])
@pytest.mark.parametrize("kind,kw", [("adam", dict(lr=1e-2, weight_decay=1e-2)), ("adamw", dict(lr=1e-2, weight_decay=5e-2))])
def test_adam_counts_steps_per_parameter_like_torch(backend, kind, kw):
    steps, world, per = 7, 2, 2
    ref, ref_steps = _branchy_reference(kind, kw, steps, world * per)
    assert sorted(set(ref_steps.values())) == [2, 7]           # `side` took part in steps 2 and 5 only
    for reload_at in (-1, 3):
        outs = run_ranks(_branchy_worker, world=world, backend=backend, args=(kind, kw, steps, per, reload_at))
        for params, counts in outs:
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-6)
            assert counts == ref_steps
