"""Per-parameter-group hyper-parameters and LR schedules (the reference applies every group's
update to every parameter, dear/dear_dopt.py:312-335, and is only correct with one group)."""
import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model


def groups_of(model):
    decay = [p for n, p in model.named_parameters() if p.dim() > 1]
    no_decay = [p for n, p in model.named_parameters() if p.dim() <= 1]
    return [dict(params=decay, weight_decay=1e-2, momentum=0.9),
            dict(params=no_decay, weight_decay=0.0, momentum=0.5, lr=0.02, nesterov=True)]


def reference(steps, world, per_rank):
    model = make_model(); model.eval()
    opt = torch.optim.SGD(groups_of(model), lr=0.05)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
        sched.step()
    return [p.detach().clone() for p in model.parameters()]


def worker(rank, world, steps, per_rank):
    import dear_pytorch_b200 as dear
    model = make_model(); model.eval()
    opt = torch.optim.SGD(groups_of(model), lr=0.05)
    opt = dear.DistributedOptimizer(opt, model, threshold=0.002, verbose=False)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
        sched.step()
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_param_groups_and_lr_schedule(backend):
    steps, world, per_rank = 6, 2, 4
    ref = reference(steps, world, per_rank)
    for params in run_ranks(worker, world=world, backend=backend, args=(steps, per_rank)):
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=3e-5, atol=3e-6)


def unused_worker(rank, world):
    """A parameter that receives no gradient: its bucket is flushed with zeros at step() and weight
    decay still applies (like torch.optim.SGD with a zero gradient would NOT — torch skips params whose
    grad is None, so the oracle treats the missing gradient as zeros only for buckets shared with others)."""
    import dear_pytorch_b200 as dear

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(6, 6)
            self.unused = nn.Linear(6, 6)
            self.b = nn.Linear(6, 3)

        def forward(self, x):
            return self.b(torch.relu(self.a(x)))
    torch.manual_seed(0)
    model = Net()
    w0 = model.unused.weight.detach().clone()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), model, threshold=None,
                                    num_nearby_layers=-1, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(3):
        x = torch.randn(4, 6, generator=torch.Generator().manual_seed(t * 10 + rank))
        opt.zero_grad()
        model(x).sum().backward()
        opt.step()
    opt.synchronize()
    return torch.equal(model.unused.weight, w0), [p.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_unused_parameters_do_not_hang_or_move(backend):
    outs = run_ranks(unused_worker, world=2, backend=backend)
    assert outs[0][0] and outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)


def frozen_worker(rank, world, steps, per):
    import dear_pytorch_b200 as dear
    from test_dear_equivalence import data, make_model
    m = make_model(); m.eval()
    for p in m[0].parameters():
        p.requires_grad_(False)
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.05, momentum=0.9, weight_decay=1e-3)
    opt = dear.DistributedOptimizer(opt, m, threshold=0.001, verbose=False)
    dear.broadcast_parameters(m.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x[rank * per:(rank + 1) * per]), y[rank * per:(rank + 1) * per]).backward()
        opt.step()
    opt.synchronize()
    return [p.detach().clone() for p in m.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_frozen_parameters_stay_out_of_the_buckets(backend):
    from test_dear_equivalence import data, make_model
    ref = make_model(); ref.eval()
    for p in ref[0].parameters():
        p.requires_grad_(False)
    opt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=0.05, momentum=0.9, weight_decay=1e-3)
    for t in range(3):
        x, y = data(t, 8)
        opt.zero_grad()
        nn.functional.cross_entropy(ref(x), y).backward()
        opt.step()
    for params in run_ranks(frozen_worker, world=2, backend=backend, args=(3, 4)):
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=3e-5, atol=3e-6)


def sched_worker(rank, world):
    import dear_pytorch_b200 as dear
    from test_dear_equivalence import data, make_model
    m = make_model(); m.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9), m, threshold=0.001, verbose=False)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)          # wraps optimizer.step like for any optimizer
    dear.broadcast_parameters(m.state_dict(), 0)
    for t in range(6):
        x, y = data(t, 4 * world)
        opt.zero_grad()
        nn.functional.cross_entropy(m(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4]).backward()
        opt.step()
        sch.step()
    opt.synchronize()
    return [p.detach().clone() for p in m.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_torch_lr_scheduler_drives_the_fused_update(backend):
    from test_dear_equivalence import data, make_model
    ref = make_model(); ref.eval()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    for t in range(6):
        x, y = data(t, 8)
        opt.zero_grad()
        nn.functional.cross_entropy(ref(x), y).backward()
        opt.step()
        sch.step()
    for params in run_ranks(sched_worker, world=2, backend=backend):
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=3e-5, atol=3e-6)
