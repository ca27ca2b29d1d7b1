"""TrainStep: the rotated loop body (update of the previous gradients -> forward -> backward) is the same
training run as the natural one once finish() / optimizer.synchronize() has applied the last update."""
import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model, reference_run


def worker(rank, world, case, steps, per_rank, overlap, flush_mid):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200.utils.train import TrainStep
    model = make_model(); model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, **case)
    opt = dear.DistributedOptimizer(opt, model, threshold=0.001, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    step = TrainStep(model, opt, nn.functional.cross_entropy, overlap_update=overlap)
    losses = []
    for t in range(steps):
        x, y = data(t, world * per_rank)
        losses.append(float(step(x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank])))
        if flush_mid and t == 1:
            opt.synchronize()                 # e.g. an evaluation or a checkpoint in the middle of training
            assert not step._pending_update
    opt.synchronize()                         # applies the deferred last update
    return losses, [p.detach().clone() for p in model.parameters()], opt.engine.num_updates


@pytest.mark.parametrize("backend", ["gloo", "emu"])
@pytest.mark.parametrize("flush_mid", [False, True])
def test_rotated_loop_equals_natural_loop(backend, flush_mid):
    case = dict(momentum=0.9, weight_decay=1e-3)
    steps, world, per_rank = 5, 2, 4
    ref = reference_run(case, steps, world, per_rank)
    natural = run_ranks(worker, world=world, backend=backend, args=(case, steps, per_rank, False, False))
    rotated = run_ranks(worker, world=world, backend=backend, args=(case, steps, per_rank, True, flush_mid))
    for (ln, pn, un), (lr, pr, ur) in zip(natural, rotated):
        assert un == ur == steps
        torch.testing.assert_close(torch.tensor(lr), torch.tensor(ln), rtol=1e-6, atol=1e-7)
        for a, b, c in zip(pr, pn, ref):
            assert torch.equal(a, b)
            torch.testing.assert_close(a, c, rtol=2e-5, atol=2e-6)


def test_rotated_loop_with_a_plain_torch_optimizer():
    from dear_pytorch_b200.utils.train import TrainStep
    outs = []
    for overlap in (False, True):
        model = make_model(); model.eval()
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        step = TrainStep(model, opt, nn.functional.cross_entropy, overlap_update=overlap)
        for t in range(4):
            step(*data(t, 8))
        step.finish()
        outs.append([p.detach().clone() for p in model.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
