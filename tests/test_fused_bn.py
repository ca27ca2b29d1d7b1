"""Fused BatchNorm(+add)(+ReLU): module semantics on CPU (composite fallback) and kernel numerics
on GPU against the plain PyTorch fp32 composite."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from dear_pytorch_b200.models.resnet import resnet18, resnet50
from dear_pytorch_b200.ops.fused_bn import FusedBatchNormAct2d, bn_act


def test_module_is_a_drop_in_batchnorm_on_cpu():
    torch.manual_seed(0)
    ref = nn.BatchNorm2d(8)
    fused = FusedBatchNormAct2d(8, relu=True)
    fused.load_state_dict(ref.state_dict())                       # identical keys
    x = torch.randn(4, 8, 5, 5, requires_grad=True)
    z = torch.randn(4, 8, 5, 5)
    y1 = F.relu(ref(x) + z)
    x2 = x.detach().clone().requires_grad_(True)
    y2 = fused(x2, residual=z)
    torch.testing.assert_close(y2, y1)
    y1.sum().backward(); y2.sum().backward()
    torch.testing.assert_close(x2.grad, x.grad)
    torch.testing.assert_close(fused.running_mean, ref.running_mean)
    assert int(fused.num_batches_tracked) == 1
    fused.eval(); ref.eval()
    torch.testing.assert_close(fused(x, residual=z), F.relu(ref(x) + z))


def test_resnet_fused_flag_keeps_parameters_and_outputs():
    torch.manual_seed(0)
    a, b = resnet18(), resnet18(fused_bn=True)
    b.load_state_dict(a.state_dict())
    assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    x = torch.randn(2, 3, 64, 64)
    torch.testing.assert_close(b(x), a(x), rtol=1e-4, atol=1e-4)
    assert sum(p.numel() for p in resnet50(fused_bn=True).parameters()) == 25557032


def _run(dtype, relu, with_res, shape):
    dev = "cuda"
    torch.manual_seed(1)
    N, C, H, W = shape
    x = (torch.randn(N, C, H, W, device=dev) * 2 + 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    z = torch.randn(N, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last) if with_res else None
    w = torch.rand(C, device=dev) + 0.5
    b = torch.randn(C, device=dev)
    dy = torch.randn(N, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    # reference in fp32
    xr = x.detach().float().clone().requires_grad_(True)       # (.float() aliases an fp32 tensor: clone)
    zr = z.detach().float().clone().requires_grad_(True) if with_res else None
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    yr = F.batch_norm(xr, rm_r, rv_r, wr, br, True, 0.1, 1e-5)
    if with_res:
        yr = yr + zr
    if relu:
        yr = F.relu(yr)
    yr.backward(dy.float())
    # ReLU is discontinuous: elements whose pre-activation is within rounding distance of zero may
    # legitimately land on the other side in the fused arithmetic -> exclude them from the comparison
    with torch.no_grad():
        pre = F.batch_norm(x.float(), None, None, w, b, True, 0.0, 1e-5) + (z.float() if with_res else 0)
        safe = (pre.abs() > (1e-4 if dtype == torch.float32 else 5e-2)) if relu else torch.ones_like(pre, dtype=torch.bool)
    # fused
    xf = x.detach().clone().requires_grad_(True)
    zf = z.detach().clone().requires_grad_(True) if with_res else None
    wf, bf = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm_f, rv_f = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    yf = bn_act(xf, wf, bf, rm_f, rv_f, True, 0.1, 1e-5, relu=relu, residual=zf)
    assert yf.is_contiguous(memory_format=torch.channels_last) and yf.dtype == dtype
    yf.backward(dy)
    tol = dict(rtol=2e-4, atol=2e-4) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(yf.float() * safe, yr * safe, **tol)
    torch.testing.assert_close(rm_f, rm_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv_f, rv_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xf.grad.float() * safe, xr.grad * safe, **tol)
    gtol = dict(rtol=2e-3, atol=2e-2) if dtype == torch.float32 else dict(rtol=5e-2, atol=0.5)
    torch.testing.assert_close(wf.grad, wr.grad, **gtol)
    torch.testing.assert_close(bf.grad, br.grad, **gtol)
    if with_res:
        torch.testing.assert_close(zf.grad.float() * safe, zr.grad * safe, **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("relu,with_res", [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize("shape", [(8, 64, 14, 14), (3, 256, 7, 5), (2, 2048, 3, 3), (5, 16, 9, 9), (2, 96, 6, 6), (3, 160, 5, 5)])
def test_fused_kernels_match_pytorch(dtype, relu, with_res, shape):
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    x = torch.empty(shape, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    if not C.bn_act_supported(x):
        pytest.skip("shape not covered by the kernel (falls back to the composite)")
    _run(dtype, relu, with_res, shape)


@pytest.mark.gpu
def test_fused_inference_path_and_fallbacks():
    dev = "cuda"
    m = FusedBatchNormAct2d(64).to(dev)
    m.running_mean.normal_(); m.running_var.uniform_(0.5, 2.0)
    ref = nn.BatchNorm2d(64).to(dev)
    ref.load_state_dict(m.state_dict())
    m.eval(); ref.eval()
    x = torch.randn(4, 64, 8, 8, device=dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        torch.testing.assert_close(m(x), F.relu(ref(x)), rtol=1e-4, atol=1e-4)
    # NCHW input and odd widths fall back to the composite
    m.train(); ref.train()
    xn = torch.randn(4, 64, 8, 8, device=dev)
    torch.testing.assert_close(m(xn), F.relu(ref(xn)), rtol=1e-4, atol=1e-4)
    odd = FusedBatchNormAct2d(24).to(dev)
    assert odd(torch.randn(2, 24, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)).shape == (2, 24, 4, 4)
