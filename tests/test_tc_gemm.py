"""tcgen05 GEMMs with fused epilogues (csrc/tc_gemm*.cu) against plain PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

from dear_pytorch_b200.ops.tc_gemm import fused_ffn, linear_bias, require_tc, tc_launches


def test_cpu_fallback_is_the_plain_formula():
    torch.manual_seed(0)
    x = torch.randn(5, 7, 16, requires_grad=True)
    w1, b1, w2, b2 = torch.randn(32, 16), torch.randn(32), torch.randn(16, 32), torch.randn(16)
    torch.testing.assert_close(fused_ffn(x, w1, b1, w2, b2), F.linear(F.gelu(F.linear(x, w1, b1)), w2, b2))
    torch.testing.assert_close(linear_bias(x, w1, b1), F.linear(x, w1, b1))


def _rand(shape, dev, scale=1.0):
    return (scale * torch.randn(shape, device=dev)).to(torch.bfloat16)


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,N", [(2048, 1024, 4096), (256, 64, 128), (300, 72, 136), (1, 8, 8), (777, 1024, 1032)])
def test_ffn_up_and_linear_bias(M, K, N):
    tc = require_tc()
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    x, w, b = _rand((M, K), dev), _rand((N, K), dev, K ** -0.5), _rand((N,), dev)
    n0 = tc_launches()
    h, z = tc.ffn_up(x, w, b)
    y = tc.linear_bias(x, w, b)
    assert tc_launches() == n0 + 2
    z_ref = x.float() @ w.float().t() + b.float()
    torch.testing.assert_close(z.float(), z_ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float(), z_ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(h.float(), F.gelu(z_ref), rtol=1e-2, atol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,N", [(2048, 1024, 4096), (256, 64, 128), (300, 72, 136), (5, 8, 16)])
def test_ffn_dgelu(M, K, N):
    tc = require_tc()
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    dy, w, z = _rand((M, K), dev), _rand((K, N), dev, K ** -0.5), _rand((M, N), dev)
    dz = tc.ffn_dgelu(dy, w, z)
    z32 = z.float().requires_grad_(True)
    F.gelu(z32).backward(dy.float() @ w.float())
    torch.testing.assert_close(dz.float(), z32.grad, rtol=1.5e-2, atol=1.5e-2)


@pytest.mark.gpu
def test_fused_ffn_autograd_matches_eager_bf16():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    H, I, M = 256, 1024, 520
    x = _rand((4, M // 4, H), dev).requires_grad_(True)
    w1, b1 = _rand((I, H), dev, H ** -0.5).requires_grad_(True), _rand((I,), dev, 0.1).requires_grad_(True)
    w2, b2 = _rand((H, I), dev, I ** -0.5).requires_grad_(True), _rand((H,), dev, 0.1).requires_grad_(True)
    dy = _rand((4, M // 4, H), dev)
    for down in (True, False):
        y = fused_ffn(x, w1, b1, w2, b2, tc_down=down)
        grads = torch.autograd.grad(y, (x, w1, b1, w2, b2), dy)
        ps = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
        y_ref = F.linear(F.gelu(F.linear(ps[0], ps[1], ps[2])), ps[3], ps[4])
        ref = torch.autograd.grad(y_ref, ps, dy.float())
        torch.testing.assert_close(y.float(), y_ref, rtol=2e-2, atol=2e-2)
        for g, r in zip(grads, ref):
            scale = r.abs().max().item()
            torch.testing.assert_close(g.float(), r, rtol=3e-2, atol=3e-2 * max(scale, 1.0))


@pytest.mark.gpu
def test_tc_ffn_inside_cuda_graph():
    dev = torch.device("cuda:0")
    tc = require_tc()
    x, w, b = _rand((512, 256), dev), _rand((1024, 256), dev, 1 / 16), _rand((1024,), dev)
    tc.ffn_up(x, w, b)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        h, z = tc.ffn_up(x, w, b)
    x.copy_(_rand((512, 256), dev))
    g.replay(); torch.cuda.synchronize()
    torch.testing.assert_close(z.float(), x.float() @ w.float().t() + b.float(), rtol=1e-2, atol=1e-2)
