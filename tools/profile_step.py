#!/usr/bin/env python
"""Device timeline of a few training steps (torch.profiler / CUPTI) and an overlap report.

    python tools/profile_step.py --model resnet50 --steps 4            # 1 GPU
    torchrun --nproc-per-node 8 tools/profile_step.py --model bert --dtype bf16

Writes (rank 0) `<out>.trace.json` (chrome://tracing / Perfetto: every kernel on every stream, the
fused kernels show up as `dear::rs_kernel` / `dear::ag_kernel` on the communication stream) and prints
  * time per step, busy time of the compute stream(s), idle gaps > 20 us on the compute stream
    (= exposed communication / launch bubbles);
  * count / mean / max duration of rs_kernel and ag_kernel (max - mean ~ time spent spinning for the
    slowest rank) and how much of their time overlaps compute kernels.
This is the measurement the reference approximates with `exclude_parts` runs (dear/batch.sh:38-42).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import dear_pytorch_b200 as dear  # noqa: E402
from dear_pytorch_b200.utils.train import TrainStep  # noqa: E402


def union_length(intervals):
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in sorted(intervals):
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def overlap_with(intervals, others):
    """Total length of `intervals` covered by the union of `others` (both lists of (s, e))."""
    others = sorted(others)
    merged = []
    for s, e in others:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    tot = 0.0
    for s, e in intervals:
        for ms, me in merged:
            if me <= s:
                continue
            if ms >= e:
                break
            tot += min(e, me) - max(s, ms)
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--threshold", type=float, default=25.0)
    ap.add_argument("--top", type=int, default=30, help="rows of the per-kernel table")
    ap.add_argument("--extra", default="", help="extra bench.py flags, e.g. '--fused-ln 0 --tc-ffn 0'")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile"))
    a = ap.parse_args()
    args = bench.parse_args(["--model", a.model] + (["--dtype", a.dtype] if a.dtype else []) +
                            (["--batch-size", str(a.batch_size)] if a.batch_size else []) +
                            ["--threshold", str(a.threshold)] + a.extra.split())
    dear.init()
    rank, world, device = dear.rank(), dear.size(), dear.device()
    torch.backends.cudnn.benchmark = True
    wl = bench.Workload(args, device, rank)
    lr = 2e-5 if wl.is_bert else 0.01 * world
    opt = dear.DistributedOptimizer(torch.optim.SGD(wl.model.parameters(), lr=lr), wl.model, threshold=args.threshold,
                                    verbose=False)
    dear.broadcast_parameters(wl.model.state_dict(), 0)
    step = TrainStep(wl.model, opt, wl.loss_fn)
    batch = wl.to_step_args(tuple(t.to(device) for t in wl.host_batches[0]))
    for _ in range(a.warmup):
        step(*batch)
    opt.engine.synchronize(host=True)
    dear.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            step(*batch)
        opt.engine.synchronize(host=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    trace = "%s.rank%d.trace.json" % (a.out, rank)
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        ev = json.load(f)["traceEvents"]
    kern = [e for e in ev if e.get("cat") == "kernel" and "dur" in e]
    if not kern:
        print("no kernel events captured (CUPTI unavailable?)")
        return
    t0 = min(e["ts"] for e in kern)
    t1 = max(e["ts"] + e["dur"] for e in kern)
    comm = [e for e in kern if "rs_kernel" in e["name"] or "ag_kernel" in e["name"] or "gen_kernel" in e["name"]]
    comp = [e for e in kern if e not in comm]
    comp_iv = [(e["ts"], e["ts"] + e["dur"]) for e in comp]
    comm_iv = [(e["ts"], e["ts"] + e["dur"]) for e in comm]
    busy = union_length(comp_iv)
    gaps = []
    prev_end = None
    for s, e in sorted(comp_iv):
        if prev_end is not None and s - prev_end > 20:
            gaps.append((s - prev_end, prev_end - t0))
        prev_end = e if prev_end is None else max(prev_end, e)
    rep = {"rank": rank, "world": world, "model": a.model, "steps": a.steps,
           "ms_per_step": (t1 - t0) / 1e3 / a.steps, "compute_busy_ms_per_step": busy / 1e3 / a.steps,
           "compute_idle_gaps_over_20us_ms_per_step": sum(g for g, _ in gaps) / 1e3 / a.steps,
           "largest_gaps_us": sorted((round(g, 1) for g, _ in gaps), reverse=True)[:8]}
    for tag in ("rs_kernel", "ag_kernel"):
        ds = [e["dur"] for e in comm if tag in e["name"]]
        iv = [(e["ts"], e["ts"] + e["dur"]) for e in comm if tag in e["name"]]
        if ds:
            rep[tag] = {"n_per_step": len(ds) / a.steps, "mean_us": sum(ds) / len(ds), "max_us": max(ds),
                        "total_ms_per_step": sum(ds) / 1e3 / a.steps,
                        "overlapped_with_compute_frac": overlap_with(iv, comp_iv) / max(sum(ds), 1e-9)}
    # per-kernel aggregate: where the GPU time of a step goes (durations are valid under CUPTI, gaps are not)
    agg = {}
    for e in kern:
        name = e["name"]
        name = name[:name.index("<")] if "<" in name and not name.startswith("void") else name
        name = name.replace("void ", "")[:100]
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += e["dur"]
    total = sum(c[1] for c in agg.values())
    rep["kernel_us_per_step"] = round(total / a.steps, 1)
    rep["top_kernels"] = [{"name": n, "per_step": round(c[0] / a.steps, 1), "us_per_step": round(c[1] / a.steps, 1),
                           "pct": round(100.0 * c[1] / total, 1)}
                          for n, c in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]]
    with open("%s.rank%d.report.json" % (a.out, rank), "w") as f:
        json.dump(rep, f, indent=1)
    if rank == 0:
        print(json.dumps(rep, indent=1))
    if rank != 0:
        os.remove(trace)          # keep the download small: one trace is enough, every rank keeps its report
    opt.engine.close()
    dear.shutdown()


if __name__ == "__main__":
    main()
