"""DeAR data-parallel SGD == single-process SGD on the concatenated batch."""

import pytest
import torch
import torch.nn as nn

from _mp import run_ranks


def make_model(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(
        nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(),
        nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.ReLU(),
        nn.AdaptiveAvgPool2d(2), nn.Flatten(),
        nn.Linear(32, 33), nn.ReLU(), nn.Linear(33, 10))


def data(step, n):
    g = torch.Generator().manual_seed(1000 + step)
    return torch.randn(n, 3, 8, 8, generator=g), torch.randint(0, 10, (n,), generator=g)


def sgd_kwargs(case):
    return dict(lr=0.05, **case)


CASES = [
    dict(),
    dict(momentum=0.9),
    dict(momentum=0.9, nesterov=True, weight_decay=1e-2),
    dict(momentum=0.8, dampening=0.3, weight_decay=5e-3),
]


def reference_run(case, steps, world, per_rank):
    model = make_model()
    # BatchNorm statistics are per-rank in data parallel; use eval-mode BN so that the
    # concatenated-batch oracle is exact.
    model.eval()
    opt = torch.optim.SGD(model.parameters(), **sgd_kwargs(case))
    for t in range(steps):
        x, y = data(t, world * per_rank)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    return [p.detach().clone() for p in model.parameters()]


def dear_worker(rank, world, case, steps, per_rank, threshold, nearby):
    import dear_pytorch_b200 as dear
    model = make_model()
    model.eval()
    opt = torch.optim.SGD(model.parameters(), **sgd_kwargs(case))
    opt = dear.DistributedOptimizer(opt, model, threshold=threshold, num_nearby_layers=nearby, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(steps):
        x, y = data(t, world * per_rank)
        x, y = x[rank * per_rank:(rank + 1) * per_rank], y[rank * per_rank:(rank + 1) * per_rank]
        opt.zero_grad()
        nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()], len(opt.engine.plan.buckets)


@pytest.mark.parametrize("backend", ["gloo", "emu"])
@pytest.mark.parametrize("case", CASES)
def test_matches_single_process_sgd(backend, case):
    steps, per_rank, world = 4, 4, 2
    ref = reference_run(case, steps, world, per_rank)
    outs = run_ranks(dear_worker, world=world, backend=backend, args=(case, steps, per_rank, 0.001, None))
    for params, nb in outs:
        assert nb > 1
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    for a, b in zip(outs[0][0], outs[1][0]):
        assert torch.equal(a, b)          # replicas are bit-identical


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_nearby_layers_and_single_bucket(backend):
    case = dict(momentum=0.9)
    ref = reference_run(case, 3, 2, 4)
    for nearby in (1, 2, -1):
        outs = run_ranks(dear_worker, world=2, backend=backend, args=(case, 3, 4, None, nearby))
        for params, nb in outs:
            for a, b in zip(params, ref):
                torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)


def test_world_of_three_emu():
    case = dict(momentum=0.9, weight_decay=1e-3)
    ref = reference_run(case, 3, 3, 2)
    outs = run_ranks(dear_worker, world=3, backend="emu", args=(case, 3, 2, 0.002, None))
    for params, nb in outs:
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)


def test_engine_on_the_emulated_pipelined_reduce_scatter():
    """The whole engine with every bucket on the stripe-pipelined Kernel A variant (host emulation), 3 ranks."""
    case = CASES[2]
    ref = reference_run(case, 4, 3, 2)
    env = {"DEAR_RS_ALGO": "pipe", "DEAR_STRIPE_MB": "0.03125"}
    from _mp import run_ranks
    for params, nb in run_ranks(dear_worker, world=3, backend="emu", args=(case, 4, 2, 0.001, None), extra_env=env):
        assert nb > 1
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)


def _direct_wgrad_worker(rank, world):
    import dear_pytorch_b200 as dear
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 4))
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), model, threshold=0.0001, verbose=False)
    eng = opt.engine
    x = torch.randn(8, 16)
    model(x).pow(2).mean().backward()
    in_bucket = {}
    for s in eng.plan.slots:
        p = s.param
        in_bucket[s.name] = (p.grad is not None and p.grad.data_ptr() == eng._grad_view[p].data_ptr(),
                             eng._src[s.bucket][s.index_in_bucket])
    opt.step()
    opt.synchronize()
    return in_bucket


def test_linear_weight_gradients_are_written_into_the_bucket_by_the_gemm():
    """True grad-as-bucket-view for GEMM-produced gradients (ops/direct_wgrad.py): after backward the Linear weights'
    ``p.grad`` aliases the gradient bucket and the pack table skips them (source pointer 0 = already in place);
    biases still travel through the pack."""
    from _mp import run_ranks
    out = run_ranks(_direct_wgrad_worker, world=2, backend="emu")[0]
    for name, (aliased, src) in out.items():
        if name.endswith("weight"):
            assert aliased and src == 0, (name, aliased, src)
        else:
            assert not aliased and src != 0, (name, aliased, src)


def test_direct_wgrad_linear_under_autocast_matches_the_stock_module():
    """autocast is off inside backward: the patched Linear must keep the operands the forward GEMM used and return
    gradients in the dtype of the fp32 masters (it used to fail with a bf16 x fp32 matmul)."""
    import copy
    from dear_pytorch_b200.ops import direct_wgrad
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4, bias=False))
    r = copy.deepcopy(m)
    assert direct_wgrad.install(m) == 2
    x = torch.randn(5, 3, 8, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    for mod, inp in ((m, x), (r, xr)):
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = mod(inp)
        assert y.dtype == torch.bfloat16
        y.float().pow(2).sum().backward()
    for a, b in zip(m.parameters(), r.parameters()):
        assert a.grad.dtype == torch.float32
        torch.testing.assert_close(a.grad, b.grad, rtol=0, atol=0)
    torch.testing.assert_close(x.grad, xr.grad, rtol=0, atol=0)


def _autocast_engine_worker(rank, world):
    import dear_pytorch_b200 as dear
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 4))
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), model, threshold=0.0001, verbose=False)
    g = torch.Generator().manual_seed(5)
    for _ in range(3):
        x = torch.randn(8, 16, generator=g)
        opt.zero_grad()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = model(x).float().pow(2).mean()
        loss.backward()
        opt.step()
    opt.synchronize()
    return [p.detach().clone() for p in model.parameters()]


def test_engine_with_autocast_forward_direct_wgrad_on_and_off():
    from _mp import run_ranks
    on = run_ranks(_autocast_engine_worker, world=2, backend="emu")[0]
    off = run_ranks(_autocast_engine_worker, world=2, backend="emu", extra_env={"DEAR_DIRECT_WGRAD": "0"})[0]
    for a, b in zip(on, off):
        torch.testing.assert_close(a, b, rtol=0, atol=0)


@pytest.mark.parametrize("algo", ["oneshot", "pipe"])
def test_world_of_eight_emu(algo):
    """The headline world size on the host emulation of the kernels (same flag protocol, same tables), both
    reduce-scatter variants: 8 ranks x 1 sample == one process on the 8-sample batch."""
    case = CASES[2]
    ref = reference_run(case, 3, 8, 1)
    env = {"DEAR_RS_ALGO": algo, "DEAR_STRIPE_MB": "0.0078125"} if algo == "pipe" else None
    outs = run_ranks(dear_worker, world=8, backend="emu", args=(case, 3, 1, 0.001, None), timeout=300, extra_env=env)
    for params, nb in outs:
        assert nb > 1
        for a, b in zip(params, ref):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    for other in outs[1:]:
        for a, b in zip(outs[0][0], other[0]):
            assert torch.equal(a, b)
