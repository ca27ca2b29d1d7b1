"""Runtime life cycle: engines can be closed (parameters handed back from the buckets), the native communicator torn
down and a new one created in the same process."""
import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model


def worker(rank, world):
    import dear_pytorch_b200 as dear
    outs = []
    for rnd in range(3):
        if rnd:
            dear.init()
        model = make_model(); model.eval()
        opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9), model, threshold=0.001,
                                        verbose=False)
        dear.broadcast_parameters(model.state_dict(), 0)
        for t in range(2):
            x, y = data(t, 4 * world)
            opt.zero_grad()
            nn.functional.cross_entropy(model(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4]).backward()
            opt.step()
        opt.synchronize()
        before = [p.detach().clone() for p in model.parameters()]
        opt.engine.close()                       # parameters become ordinary tensors again
        assert opt.engine.backend is None
        for p, b in zip(model.parameters(), before):
            assert torch.equal(p, b) and p.grad is None
        with torch.no_grad():
            model(data(0, 2)[0])                 # the model is usable without the engine
        outs.append(float(sum(float(p.detach().abs().sum()) for p in model.parameters())))
        dear.shutdown(destroy_process_group=False)
        assert not dear.is_initialized()
    dear.init()                                  # the harness shuts down once more after the worker returns
    return outs


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_close_shutdown_and_reinitialise_in_one_process(backend):
    outs = run_ranks(worker, world=2, backend=backend)
    assert outs[0] == outs[1] and len(set(outs[0])) == 1
