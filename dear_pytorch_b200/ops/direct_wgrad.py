"""Weight gradients of ``nn.Linear`` written by the GEMM straight into the symmetric gradient bucket.

SURVEY §2.2 K3 asks for "grad-as-bucket-view": the reference copies every gradient into its fusion buffer with one
``copy_`` per parameter (dear/dear_dopt.py:265), and Kernel A's pack phase is the fused form of that copy (135 us of a
740 us reduce-scatter for VGG-16's 392 MB fc bucket at 8 GPUs, round-1 profile).  Autograd does not let a layer choose
where its gradient is allocated — but for a Linear layer the gradient IS a GEMM, and a GEMM can write anywhere:

    dW = dY^T X      ->      torch.mm(dY^T, X, out=<the parameter's slice of the gradient bucket>)

The backward below does exactly that when the DeAR engine has attached a bucket view to the weight
(``weight._dear_grad_view``, set by ``DearEngine._build`` on the fused backend), and hands autograd a fresh alias of
that view: AccumulateGrad adopts it without a copy, the engine's hook sees ``p.grad`` already inside the bucket and drops
the parameter from Kernel A's pack table.  For VGG-16 this removes 472 of 553 MB of pack traffic per step, for BERT
~97 % of it.  Convolution weight gradients come out of cuDNN into fresh tensors and still go through the pack.

``install(model)`` re-routes the forward of every plain ``nn.Linear`` of a model through this function (idempotent;
``DEAR_DIRECT_WGRAD=0`` disables it).
"""
from __future__ import annotations

import os
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class _LinearDirectWgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.has_bias = bias is not None
        ctx.grad_dtypes = (x.dtype, weight.dtype, bias.dtype if bias is not None else None)
        dev = x.device.type
        if torch.is_autocast_enabled(dev) and x.is_floating_point():
            # autocast is off inside backward: keep the operands in the dtype the forward GEMM actually used, and
            # hand gradients back in the dtype of the tensors autograd knows (fp32 master weights under bf16 autocast)
            dt = torch.get_autocast_dtype(dev)
            x, weight = x.to(dt), weight.to(dt)
            bias = bias.to(dt) if bias is not None else None
            with torch.autocast(dev, enabled=False):
                y = F.linear(x, weight, bias)
        else:
            y = F.linear(x, weight, bias)
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors            # as the GEMM saw them (cast copies under autocast)
        xdt, wdt, bdt = ctx.grad_dtypes
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = dy2.mm(weight).view(x.shape).to(xdt)
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, x.shape[-1])
            # under autocast `weight` is a temporary copy without the engine's attributes: ordinary gradient there
            view = getattr(weight, "_dear_grad_view", None)
            if (view is not None and not getattr(weight, "_dear_grad_written", False) and view.dtype == dy2.dtype
                    and x2.dtype == dy2.dtype and view.is_contiguous() and view.device == dy2.device):
                torch.mm(dy2.t(), x2, out=view)          # the GEMM's epilogue writes the bucket
                dw = view.detach()                       # fresh alias: AccumulateGrad adopts it without a copy
                # a weight used twice in one forward (tied layers) gets ONE direct write; the other uses produce
                # ordinary tensors that autograd sums into it.  The engine clears the mark at step().
                weight._dear_grad_written = True
            else:
                dw = dy2.t().mm(x2).to(wdt)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0).to(bdt)
        return dx, dw, db


def linear(x, weight, bias=None):
    """``F.linear`` whose weight gradient lands in ``weight._dear_grad_view`` when the engine attached one."""
    return _LinearDirectWgrad.apply(x, weight, bias)


def _forward(self, x):
    return _LinearDirectWgrad.apply(x, self.weight, self.bias)


def enabled() -> bool:
    return os.environ.get("DEAR_DIRECT_WGRAD", "1") not in ("0", "false", "False")


def install(model: nn.Module) -> int:
    """Route every plain ``nn.Linear`` of ``model`` through the direct-wgrad function; returns how many were patched."""
    n = 0
    for m in model.modules():
        if type(m) is nn.Linear and not getattr(m, "_dear_direct_wgrad", False):
            m.forward = types.MethodType(_forward, m)
            m._dear_direct_wgrad = True
            n += 1
    return n
