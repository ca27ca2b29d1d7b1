"""Skewed ranks, many tiny buckets, shared parameters, execution order != registration order, eval-mode
forwards between training steps (protocol robustness; §8.2 ordering rules)."""
import random

import pytest
import time

import torch
import torch.nn as nn

from _mp import run_ranks


class Tangled(nn.Module):
    """Registered: head, blocks, stem, tied.  Executed: stem, blocks in reverse, tied (= stem weight), head."""

    def __init__(self, depth=10, width=24):
        super().__init__()
        torch.manual_seed(11)
        self.head = nn.Linear(width, 5)
        self.blocks = nn.ModuleList(nn.Linear(width, width) for _ in range(depth))
        self.stem = nn.Linear(width, width)
        self.unused = nn.Linear(3, 3)                    # never executed: its gradient stays absent
        self.tied = nn.Linear(width, width, bias=False)
        self.tied.weight = self.stem.weight              # shared parameter

    def forward(self, x):
        x = torch.tanh(self.stem(x))
        for b in reversed(self.blocks):
            x = x + 0.1 * torch.tanh(b(x))
        return self.head(torch.tanh(self.tied(x)))


def batch(step, n):
    g = torch.Generator().manual_seed(500 + step)
    return torch.randn(n, 24, generator=g), torch.randint(0, 5, (n,), generator=g)


def oracle(steps, world, per_rank):
    m = Tangled()
    opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-3)
    for t in range(steps):
        x, y = batch(t, world * per_rank)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x), y).backward()
        opt.step()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def worker(rank, world, steps, per_rank, nearby, jitter):
    import dear_pytorch_b200 as dear
    rnd = random.Random(1234 + rank)
    m = Tangled()
    opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-3)
    opt = dear.DistributedOptimizer(opt, m, threshold=None, num_nearby_layers=nearby, verbose=False)
    dear.broadcast_parameters(m.state_dict(), 0)

    def nap():
        if jitter:
            time.sleep(rnd.random() * jitter)
    for t in range(steps):
        x, y = batch(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
        opt.zero_grad()
        nap()
        loss = nn.functional.cross_entropy(m(x), y)
        nap()
        loss.backward()
        nap()
        opt.step()
        if t % 3 == 1:                                   # an evaluation pass between two training steps
            with torch.no_grad():
                m.eval(); m(x); m.train()
    opt.synchronize()                                    # last-step flush
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, len(opt.engine.plan.buckets)


def check(outs, ref):
    for sd, _ in outs:
        for k, v in ref.items():
            if k.startswith("unused."):
                # a gradient absent on every rank is reduced as zeros: weight decay still applies here, whereas
                # torch.optim.SGD skips parameters whose .grad is None (documented in DESIGN.md)
                continue
            torch.testing.assert_close(sd[k], v, rtol=5e-5, atol=5e-6, msg=lambda s, k=k: "%s: %s" % (k, s))
    for k in ref:
        assert all(torch.equal(outs[0][0][k], o[0][k]) for o in outs[1:]), k


def test_tangled_model_per_layer_buckets_gloo():
    ref = oracle(5, 2, 3)
    outs = run_ranks(worker, world=2, backend="gloo", args=(5, 3, 1, 0.0))
    assert outs[0][1] >= 10
    check(outs, ref)


def test_skewed_ranks_many_buckets_emu():
    """4 ranks, one bucket per layer, random stalls of up to 3 ms at every phase on every rank: the
    flag protocol (epochs, per-channel signal pads) must neither deadlock nor mix iterations."""
    steps = 12
    ref = oracle(steps, 4, 2)
    outs = run_ranks(worker, world=4, backend="emu", args=(steps, 2, 1, 0.003), timeout=240)
    assert outs[0][1] >= 10
    check(outs, ref)


def test_unused_layer_and_single_bucket_emu():
    ref = oracle(4, 3, 2)
    outs = run_ranks(worker, world=3, backend="emu", args=(4, 2, -1, 0.001))
    assert outs[0][1] == 1
    check(outs, ref)


def _skip_worker(rank, world):
    """A parameter without a gradient is skipped like torch.optim does (no weight decay, no momentum decay)."""
    import torch
    import torch.nn as nn
    import dear_pytorch_b200 as dear

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.used = nn.Linear(8, 8)
            self.unused = nn.Linear(8, 8)

        def forward(self, x):
            return self.used(x)

    torch.manual_seed(0)
    model, ref = Net(), Net()
    ref.load_state_dict(model.state_dict())
    kw = dict(lr=0.1, weight_decay=0.1, momentum=0.9)
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), **kw), model, threshold=None, num_nearby_layers=1,
                                    verbose=False)
    ropt = torch.optim.SGD(ref.parameters(), **kw)
    g = torch.Generator().manual_seed(5)
    for _ in range(3):
        x = torch.randn(world * 4, 8, generator=g)
        opt.zero_grad()
        model(x[rank * 4:(rank + 1) * 4]).pow(2).mean().backward()
        opt.step()
        ropt.zero_grad()
        ref(x).pow(2).mean().backward()
        ropt.step()
    opt.synchronize()
    return [(n, p.detach().clone(), dict(ref.named_parameters())[n].detach().clone()) for n, p in model.named_parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_parameters_without_gradient_are_not_updated(backend):
    from _mp import run_ranks
    for out in run_ranks(_skip_worker, world=2, backend=backend):
        for n, a, b in out:
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=lambda m: "%s: %s" % (n, m))


def _late_start_worker(rank, world, steps):
    import dear_pytorch_b200 as dear
    from test_adam import _Branchy, _branchy_data
    m = _Branchy()
    opt = dear.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.8, dampening=0.3, weight_decay=5e-3), m,
                                    threshold=0.0005, verbose=False)
    dear.broadcast_parameters(m.state_dict(), 0)
    for t in range(steps):
        x, y = _branchy_data(t, 4)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x[rank * 2:(rank + 1) * 2], t % 3 == 2), y[rank * 2:(rank + 1) * 2]).backward()
        opt.step()
        if t == 0:
            opt.load_state_dict(opt.state_dict())       # "never had a gradient" survives a state-dict round trip
    opt.synchronize()
    return [p.detach().clone() for p in m.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_momentum_buffer_of_a_late_starting_parameter_with_dampening(backend):
    """torch.optim.SGD initialises a momentum buffer with the parameter's FIRST gradient, undampened — also when that
    gradient arrives in step 2 (the side branch) rather than in the global first step."""
    from test_adam import _Branchy, _branchy_data
    steps = 7
    ref = _Branchy()
    opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.8, dampening=0.3, weight_decay=5e-3)
    for t in range(steps):
        x, y = _branchy_data(t, 4)
        opt.zero_grad()
        nn.functional.cross_entropy(ref(x, t % 3 == 2), y).backward()
        opt.step()
    for params in run_ranks(_late_start_worker, world=2, backend=backend, args=(steps,)):
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=3e-5, atol=3e-6)
