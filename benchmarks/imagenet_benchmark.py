#!/usr/bin/env python
"""Synthetic-ImageNet training benchmark (counterpart of */imagenet_benchmark.py in the reference).

    torchrun --nproc-per-node 8 benchmarks/imagenet_benchmark.py --model resnet50 --batch-size 64 --method dear
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from common import dear  # noqa: E402
from dear_pytorch_b200.models.registry import create, input_size  # noqa: E402
from dear_pytorch_b200.utils.train import TrainStep  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description="DeAR synthetic ImageNet benchmark",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--model", type=str, default="resnet50")
    ap.add_argument("--channels-last", type=int, default=1)
    ap.add_argument("--image-size", type=int, default=None, help="default: the model's native resolution (224; 299 for inception)")
    ap.add_argument("--fused-bn", type=int, default=1, help="fused BatchNorm(+add)+ReLU kernels (resnet*/densenet*)")
    common.add_common_args(ap)
    args = ap.parse_args(argv)
    method, cuda = common.init_runtime(args)
    device = dear.device()
    dtype = args.dtype or ("bf16" if args.fp16 else "fp32")

    kw = {"fused_bn": True} if (args.fused_bn and args.channels_last and args.model.startswith(("resnet", "densenet"))) else {}
    model = create(args.model, **kw).to(device)
    if args.channels_last and cuda:
        model = model.to(memory_format=torch.channels_last)
    if dtype == "bf16":
        model = model.to(torch.bfloat16)
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.float()
    size = args.image_size or input_size(args.model)
    data = torch.randn(args.batch_size, 3, size, size, device=device)
    if dtype == "bf16":
        data = data.to(torch.bfloat16)
    if args.channels_last and cuda:
        data = data.contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, 1000, (args.batch_size,), device=device)

    lr_scaler = 1 if args.use_adasum else dear.size()          # reference: dear/imagenet_benchmark.py:85
    optimizer = common.make_base_optimizer(args, model.parameters(), 0.01 * lr_scaler)

    def profile():
        from dear_pytorch_b200.utils.profiling import benchmark
        return benchmark(model, (data, target), lambda o, t: F.cross_entropy(o.float(), t), task="imagenet",
                         warmup=3, iters=10)
    model, optimizer = common.wrap_optimizer(method, args, model, optimizer, profile)
    if dear.size() > 1 and method != "single":
        dear.broadcast_parameters(model.state_dict(), root_rank=0)

    step = TrainStep(model, optimizer, lambda o, t: F.cross_entropy(o.float(), t),
                     autocast_dtype=torch.bfloat16 if dtype == "amp" else None, use_graph=bool(args.graph) and cuda)

    def sync(host=True):
        if hasattr(optimizer, "_dear"):
            optimizer._dear.synchronize(host=host)
        elif hasattr(optimizer, "synchronize") and method in ("dear-rb", "bytescheduler"):
            optimizer.synchronize()            # bytescheduler: deferred per-layer updates + its scheduler thread
        if cuda and host:
            torch.cuda.synchronize()

    common.log("Model: %s" % args.model)
    common.log("Method: %s, dtype: %s, backend: %s" % (method, dtype, dear.backend()))
    common.log("Batch size: %d" % args.batch_size)
    common.log("Number of %ss: %d" % ("GPU" if cuda else "CPU", dear.size()))
    res = common.run_timing(lambda: step(data, target), args, "img", args.batch_size, sync)
    common.finish(args, res, {"model": args.model, "method": method, "dtype": dtype, "world": dear.size(),
                              "batch_size": args.batch_size})
    dear.shutdown()
    return res


if __name__ == "__main__":
    main()
