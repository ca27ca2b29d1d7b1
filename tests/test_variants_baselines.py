"""Every optimizer variant and NCCL-style baseline is equivalent to single-process SGD."""
import pytest
import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model, reference_run

CASE = dict(momentum=0.9, weight_decay=1e-3)


def worker(rank, world, kind, steps, per_rank):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200.parallel import variants
    from dear_pytorch_b200.parallel.baselines import WFBPDistributedOptimizer, wrap_ddp
    from dear_pytorch_b200.parallel.baselines.horovod_like import ByteSchedulerLikeOptimizer, HorovodLikeOptimizer
    from dear_pytorch_b200.utils.profiling import benchmark
    model = make_model()
    model.eval()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, **CASE)
    fwd = model
    if kind == "naive":
        opt = variants.NaiveDistributedOptimizer(opt, model, verbose=False)
    elif kind == "wt":
        opt = variants.WaitTimeDistributedOptimizer(opt, model, cycle_time_ms=0.05, warmup_steps=2, verbose=False)
    elif kind == "rb":
        opt = variants.ReduceBroadcastDistributedOptimizer(opt, model, threshold=0.002, verbose=False)
    elif kind == "wfbp":
        opt = WFBPDistributedOptimizer(opt, model=model, threshold=0, verbose=False)
    elif kind == "wfbp-threshold":
        opt = WFBPDistributedOptimizer(opt, model=model, threshold=600, verbose=False)
    elif kind in ("mgwfbp", "asc"):
        x, y = data(0, 4)
        seq, times, sizes = benchmark(model, (x, y), nn.functional.cross_entropy, warmup=1, iters=2)
        seq, times = dear.runtime.broadcast_object((seq, times), src=0)
        opt = WFBPDistributedOptimizer(opt, model=model, seq_layernames=seq, layerwise_times=times, mgwfbp=(kind == "mgwfbp"),
                                       asc=(kind == "asc"), alpha=1e-4, beta=1e-9, verbose=False)
    elif kind == "horovod":
        opt = HorovodLikeOptimizer(opt, model, verbose=False)
    elif kind == "bytescheduler":
        opt = ByteSchedulerLikeOptimizer(opt, model, partition_mb=0.001, verbose=False)
    elif kind in ("ddp", "ddp-zero"):
        fwd, opt = wrap_ddp(model, torch.optim.SGD, dict(lr=0.05, **CASE), zero=(kind == "ddp-zero"))
    if kind not in ("ddp", "ddp-zero"):
        dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
        opt.zero_grad()
        nn.functional.cross_entropy(fwd(x), y).backward()
        opt.step()
    if hasattr(opt, "synchronize") and kind in ("naive", "wt", "rb"):
        opt.synchronize()
    info = None
    if kind == "wt":
        info = (opt.wait_time.done, len(opt.engine.plan.buckets))
    return [p.detach().clone() for p in model.parameters()], info


@pytest.mark.parametrize("kind", ["naive", "wt", "rb", "wfbp", "wfbp-threshold", "mgwfbp", "asc", "horovod",
                                  "bytescheduler", "ddp", "ddp-zero"])
def test_variant_matches_sgd(kind):
    steps, per_rank, world = 5, 3, 2
    ref = reference_run(CASE, steps, world, per_rank)
    outs = run_ranks(worker, world=world, backend="gloo", args=(kind, steps, per_rank))
    for params, info in outs:
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=3e-5, atol=3e-6)
        if kind == "wt":
            assert info[0] and info[1] >= 1


def test_sparse_topk_allgather_path_runs():
    def sparse_worker(rank, world):
        import dear_pytorch_b200 as dear
        from dear_pytorch_b200.parallel.baselines import WFBPDistributedOptimizer
        model = make_model()
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        opt = WFBPDistributedOptimizer(opt, model=model, compression="topk", is_sparse=True, density=0.25, threshold=0,
                                       verbose=False)
        dear.broadcast_parameters(model.state_dict(), 0)
        losses = []
        for t in range(6):
            x, y = data(0, 8)
            opt.zero_grad()
            loss = nn.functional.cross_entropy(model(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4])
            loss.backward()
            opt.step()
            losses.append(float(loss))
        return losses, [p.detach().clone() for p in model.parameters()]
    outs = run_ranks(sparse_worker, world=2, backend="gloo")
    assert all(torch.equal(a, b) for a, b in zip(outs[0][1], outs[1][1]))
    assert outs[0][0][-1] < outs[0][0][0] + 0.5
