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


def _sched_worker(rank, world, overlap, steps):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200.utils.train import TrainStep
    model = make_model(); model.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9), model, threshold=0.001, verbose=False)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.7)          # the LR changes after EVERY step
    dear.broadcast_parameters(model.state_dict(), 0)
    step = TrainStep(model, opt, nn.functional.cross_entropy, overlap_update=overlap)
    losses = []
    for t in range(steps):
        x, y = data(t, world * 2)
        losses.append(float(step(x[rank * 2:(rank + 1) * 2], y[rank * 2:(rank + 1) * 2])))
        sched.step()
        if t == 2:
            opt.synchronize()                # applies the pending update with the LR of step 2, not the scheduler's new one
    opt.synchronize()
    return losses, [p.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_rotated_loop_with_a_per_step_lr_schedule_matches_torch(backend):
    """The rotated body applies the update of call t at the start of call t+1 — after the user's ``scheduler.step()``.
    It must still use call t's learning rate (``DearEngine.freeze_hyper``): same result as the plain
    ``optimizer.step(); scheduler.step()`` loop of torch."""
    steps, world = 6, 2
    model = make_model(); model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.7)
    for t in range(steps):
        x, y = data(t, world * 2)
        opt.zero_grad(); nn.functional.cross_entropy(model(x), y).backward(); opt.step(); sched.step()
    natural = run_ranks(_sched_worker, world=world, backend=backend, args=(False, steps))
    rotated = run_ranks(_sched_worker, world=world, backend=backend, args=(True, steps))
    for (ln, pn), (lr, pr) in zip(natural, rotated):
        assert ln == lr
        for a, b, c in zip(pr, pn, model.parameters()):
            assert torch.equal(a, b)
            torch.testing.assert_close(a, c.detach(), rtol=2e-5, atol=2e-6)


def test_rotated_loop_with_a_plain_torch_optimizer_and_scheduler():
    from dear_pytorch_b200.utils.train import TrainStep
    outs = []
    for overlap in (False, True):
        model = make_model(); model.eval()
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
        step = TrainStep(model, opt, nn.functional.cross_entropy, overlap_update=overlap)
        for t in range(4):
            step(*data(t, 8))
            sched.step()
        step.finish()
        assert opt.param_groups[0]["lr"] == pytest.approx(0.05 * 0.5 ** 4)        # the live schedule is untouched
        outs.append([p.detach().clone() for p in model.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
