"""Checkpoint / resume of the sharded state, incl. resuming into a different bucket layout and
loading a checkpoint into stock torch.optim.SGD."""
import os
import tempfile

import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model, reference_run

CASE = dict(momentum=0.9, weight_decay=1e-3)


def worker(rank, world, path, steps_a, steps_b, per_rank):
    import dear_pytorch_b200 as dear

    def train(model, opt, t0, t1):
        for t in range(t0, t1):
            x, y = data(t, world * per_rank)
            x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
            opt.zero_grad()
            nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
    model = make_model(); model.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, **CASE), model, threshold=0.001, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    train(model, opt, 0, steps_a)
    dear.save_checkpoint(path, model, opt, extra={"step": steps_a})
    opt.engine.close()

    # resume into a fresh model with a DIFFERENT bucket layout
    model2 = make_model(seed=123); model2.eval()
    opt2 = dear.DistributedOptimizer(torch.optim.SGD(model2.parameters(), lr=0.5), model2, threshold=None, num_nearby_layers=-1,
                                     verbose=False)
    extra = dear.load_checkpoint(path, model2, opt2)
    assert extra["step"] == steps_a
    assert opt2.param_groups[0]["lr"] == 0.05 and opt2.param_groups[0]["momentum"] == 0.9
    train(model2, opt2, steps_a, steps_a + steps_b)
    opt2.synchronize()
    return [p.detach().clone() for p in model2.parameters()]


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_save_resume_matches_uninterrupted_training(backend):
    steps_a, steps_b, per_rank, world = 3, 3, 4, 2
    ref = reference_run(CASE, steps_a + steps_b, world, per_rank)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt.pt")
        outs = run_ranks(worker, world=world, backend=backend, args=(path, steps_a, steps_b, per_rank))
        for params in outs:
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=3e-5, atol=3e-6)
        # the optimizer part of the checkpoint is a stock torch.optim.SGD state dict
        ckpt = torch.load(path, weights_only=False)
        m = make_model()
        sgd = torch.optim.SGD(m.parameters(), lr=0.05, **CASE)
        sd = {"state": {k: {"momentum_buffer": v["momentum_buffer"]} for k, v in ckpt["optimizer"]["state"].items()},
              "param_groups": ckpt["optimizer"]["param_groups"]}
        sgd.load_state_dict(sd)
        assert len(sgd.state) == len(list(m.parameters()))


def save_worker(rank, world, path, steps, per_rank):
    import dear_pytorch_b200 as dear
    model = make_model(); model.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, **CASE), model, threshold=0.001, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x[rank * per_rank:(rank + 1) * per_rank]), y[rank * per_rank:(rank + 1) * per_rank]).backward()
        opt.step()
    dear.save_checkpoint(path, model, opt)
    return True


def resume_worker(rank, world, path, t0, steps, global_batch):
    import dear_pytorch_b200 as dear
    per_rank = global_batch // world
    model = make_model(seed=77); model.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.5), model, threshold=0.002, verbose=False)
    dear.load_checkpoint(path, model, opt)
    for t in range(t0, t0 + steps):
        x, y = data(t, global_batch)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x[rank * per_rank:(rank + 1) * per_rank]), y[rank * per_rank:(rank + 1) * per_rank]).backward()
        opt.step()
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()]


def test_resume_on_a_different_world_size():
    """Momentum and parameters are saved per parameter name, so a checkpoint written by 2 ranks resumes on 4 (and the
    sharding 1/2 -> 1/4 changes underneath); same global batch, so the run equals the uninterrupted one."""
    global_batch, steps_a, steps_b = 8, 3, 3
    ref = reference_run(CASE, steps_a + steps_b, 2, global_batch // 2)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt.pt")
        run_ranks(save_worker, world=2, backend="emu", args=(path, steps_a, global_batch // 2))
        for params in run_ranks(resume_worker, world=4, backend="emu", args=(path, steps_a, steps_b, global_batch)):
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=3e-5, atol=3e-6)
