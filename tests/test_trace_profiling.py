"""Auxiliary subsystems: the chrome-trace timeline wired through DEAR_TIMELINE and the
layer-wise profiler that feeds the MG-WFBP / wait-time planners."""
import json
import os
import tempfile

import torch
import torch.nn as nn

from _mp import run_ranks
from test_dear_equivalence import data, make_model


def traced_worker(rank, world, path):
    import dear_pytorch_b200 as dear
    model = make_model(); model.eval()
    opt = dear.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05), model, threshold=0.001, verbose=False)
    dear.broadcast_parameters(model.state_dict(), 0)
    for t in range(3):
        x, y = data(t, 4 * world)
        opt.zero_grad()
        nn.functional.cross_entropy(model(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4]).backward()
        opt.step()
    opt.synchronize()
    nb = len(opt.engine.plan.buckets)
    opt.engine.close()
    return nb


def test_timeline_is_written_per_rank_and_is_valid_chrome_trace():
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "trace.json")
        outs = run_ranks(traced_worker, world=2, backend="gloo", args=(path,), extra_env={"DEAR_TIMELINE": path})
        files = sorted(f for f in os.listdir(d) if f.startswith("trace"))
        assert files, "no timeline written"
        for f in files:
            with open(os.path.join(d, f)) as fh:
                events = json.load(fh)
            names = {e["name"] for e in events}
            assert any("reduce" in n.lower() or "rs" in n.lower() for n in names), names
            assert any("gather" in n.lower() or "ag" in n.lower() for n in names), names
            rows = {e["tid"] for e in events}
            assert len(rows) >= outs[0]                  # one row per bucket
            opened = {}
            for e in events:                              # B/E events are balanced per (row, name)
                k = (e["tid"], e["name"])
                if e["ph"] == "B":
                    opened[k] = opened.get(k, 0) + 1
                elif e["ph"] == "E":
                    opened[k] = opened.get(k, 0) - 1
            assert all(v == 0 for v in opened.values()), opened


def test_layerwise_profiler_reports_every_parameterised_layer():
    from dear_pytorch_b200.utils.profiling import benchmark
    model = make_model()
    x, y = data(0, 8)
    names, times, sizes = benchmark(model, (x, y), nn.CrossEntropyLoss(), task="imagenet", warmup=1, iters=2)[:3]
    n_layers = sum(1 for m in model.modules() if any(True for _ in m.parameters(recurse=False)))
    assert len(names) == len(times) == len(sizes) >= n_layers - 1
    assert all(t >= 0 for t in times) and sum(sizes) == sum(p.numel() for p in model.parameters())
