"""BO tuner behaviour and re-bucketing at the safe point."""
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model, reference_run


def test_tuner_converges_on_convex_cost():
    from dear_pytorch_b200.parallel.tuner import Tuner
    f = lambda x: 0.050 + 1e-6 * (x - 60.0) ** 2          # iteration time (s) vs threshold (MB)
    now = [0.0]
    cur = [25.0]

    def clock():
        now[0] += f(cur[0])
        return now[0]
    t = Tuner(x=25.0, bound=(1.0, 256.0), max_num_steps=10, interval=5, clock=clock, verbose=False)
    applied = []
    for _ in range(200):
        nxt = t.step()
        if nxt is not None:
            assert 1.0 <= nxt <= 256.0
            cur[0] = nxt
            applied.append(nxt)
        if t.finished:
            break
    assert t.finished and len(t.history) == 10
    best, best_time = t.opt_point()
    assert best_time <= f(25.0) + 1e-12                     # never worse than the starting point
    assert abs(best - 60.0) < 45.0                          # moved towards the optimum
    assert cur[0] == best                                   # ends on the best point
    assert t.step() is None                                 # and stays quiet afterwards


def rebucket_worker(rank, world, case, steps, per_rank, switch_at):
    import dear_pytorch_b200 as dear
    model = make_model()
    model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, **case)
    opt = dear.DistributedOptimizer(opt, model, threshold=0.001, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    nb = [len(opt.engine.plan.buckets)]
    for t in range(steps):
        if t == switch_at:
            opt.engine.request_rebucket(("threshold", 0.004))
        if t == switch_at + 2:
            opt.engine.request_rebucket(("nearby", -1))
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
        nb.append(len(opt.engine.plan.buckets))
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()], nb


def test_rebucket_preserves_training_state():
    case = dict(momentum=0.9, weight_decay=1e-3)
    steps = 7
    ref = reference_run(case, steps, 2, 4)
    for backend in ("gloo", "emu"):
        outs = run_ranks(rebucket_worker, world=2, backend=backend, args=(case, steps, 4, 2))
        for params, nb in outs:
            assert len(set(nb)) == 3, nb                    # three different layouts were used
            assert nb[-1] == 1
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)


def bo_worker(rank, world):
    import dear_pytorch_b200 as dear
    model = make_model()
    model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    opt = dear.DistributedOptimizer(opt, model, threshold=0.002, verbose=False, bo_tuning=True,
                                    bo_kwargs=dict(bound=(0.0005, 0.02), max_num_steps=3, interval=5))
    dear.broadcast_parameters(model.state_dict(), 0)
    seen = set()
    for t in range(40):
        x, y = data(t, 2 * world)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x[rank * 2:(rank + 1) * 2]), y[rank * 2:(rank + 1) * 2]).backward()
        opt.step()
        seen.add(opt.engine.plan.policy)
    opt.synchronize()
    return sorted(seen), opt.tuner.finished, [p.detach().clone() for p in model.parameters()]


def test_bo_tuning_end_to_end_is_rank_consistent():
    outs = run_ranks(bo_worker, world=2, backend="gloo")
    assert outs[0][0] == outs[1][0]                          # same sequence of policies on every rank
    assert len(outs[0][0]) >= 2 and outs[0][1] and outs[1][1]
    for a, b in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, b)
