"""Fused dropout + add + LayerNorm (csrc/ln_fused.cu) against a plain PyTorch fp32 reference."""
import pytest
import torch
import torch.nn.functional as F

from dear_pytorch_b200.ops import native
from dear_pytorch_b200.ops.fused_ln import FusedDropoutAddLayerNorm, dropout_add_layer_norm


def test_cpu_composite_and_state_dict_keys():
    torch.manual_seed(0)
    m = FusedDropoutAddLayerNorm(48, eps=1e-12, p=0.1)
    ref = torch.nn.LayerNorm(48, eps=1e-12)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(m.state_dict())
    m.eval()
    a, r = torch.randn(3, 5, 48), torch.randn(3, 5, 48)
    torch.testing.assert_close(m(a, r), ref(a + r))
    m.train()
    torch.manual_seed(1); y1 = m(a, r)
    torch.manual_seed(1); y2 = ref(r + F.dropout(a, 0.1, True))
    torch.testing.assert_close(y1, y2)


def _reference(a, r, w, b, keep, p, eps):
    """fp32 math on the values the kernel sees; s is rounded to the storage dtype like the kernel does."""
    scale = 1.0 / (1.0 - p) if keep is not None else 1.0
    da = a.float() * (keep.float() * scale if keep is not None else 1.0)
    s = (r.float() + da).to(a.dtype).float()
    return F.layer_norm(s, (a.shape[-1],), w.float(), b.float(), eps), s


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2048, 1024), (4, 7, 768), (3, 64), (515, 136), (1, 8)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_forward_backward_match_reference(dtype, shape, p):
    if dtype == torch.float32 and shape[-1] % 4:
        pytest.skip("fp32 needs H % 4 == 0")
    C = native()
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    a = torch.randn(shape, device=dev).to(dtype)
    r = torch.randn(shape, device=dev).to(dtype)
    w = (1.0 + 0.1 * torch.randn(shape[-1], device=dev)).to(dtype)
    b = (0.1 * torch.randn(shape[-1], device=dev)).to(dtype)
    eps = 1e-12 if dtype == torch.float32 else 1e-5
    assert C.ln_supported(a)
    y, s, mean, rstd, mask = C.ln_forward(a, r, w, b, p, True, eps)
    keep = mask.bool() if p > 0 else None
    if p > 0:
        frac = keep.float().mean().item()
        if a.numel() >= 4096:
            assert abs(frac - (1 - p)) < 0.02, frac
    y_ref, s_ref = _reference(a, r, w, b, keep, p, eps)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    stol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=8e-3, atol=1e-2)
    torch.testing.assert_close(s.float(), s_ref, **stol)      # (fma contraction may move one rounding)
    y_ref = F.layer_norm(s.float(), (shape[-1],), w.float(), b.float(), eps)
    torch.testing.assert_close(y.float(), y_ref, **tol)
    torch.testing.assert_close(mean, s.float().reshape(-1, shape[-1]).mean(-1), rtol=1e-5, atol=1e-5)
    # backward against autograd of the fp32 reference
    dy = torch.randn(shape, device=dev).to(dtype)
    a32 = a.float().requires_grad_(True); r32 = r.float().requires_grad_(True)
    w32 = w.float().requires_grad_(True); b32 = b.float().requires_grad_(True)
    scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    s32 = r32 + a32 * (keep.float() * scale if keep is not None else 1.0)
    # normalise the rounded s (what the kernel stored) but keep the graph: straight-through rounding
    s32r = s32 + (s.float() - s32).detach()
    F.layer_norm(s32r, (shape[-1],), w32, b32, eps).backward(dy.float())
    d_res, d_a, dgamma, dbeta, _ = C.ln_backward(dy, s, mean, rstd, w, mask, p)
    btol = dict(rtol=2e-4, atol=2e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(d_res.float(), r32.grad, **btol)
    torch.testing.assert_close(d_a.float(), a32.grad, **btol)
    rows = a.numel() // shape[-1]
    gtol = dict(rtol=1e-3, atol=1e-3 * max(1.0, rows ** 0.5)) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2 * max(1.0, rows ** 0.5))
    torch.testing.assert_close(dgamma.float(), w32.grad, **gtol)
    torch.testing.assert_close(dbeta.float(), b32.grad, **gtol)


@pytest.mark.gpu
def test_autograd_function_and_fresh_masks():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = FusedDropoutAddLayerNorm(256, eps=1e-5, p=0.0).to(dev)
    ref = torch.nn.LayerNorm(256, eps=1e-5).to(dev)
    ref.load_state_dict(m.state_dict())
    a = torch.randn(33, 256, device=dev, requires_grad=True)
    r = torch.randn(33, 256, device=dev, requires_grad=True)
    a2, r2 = a.detach().clone().requires_grad_(True), r.detach().clone().requires_grad_(True)
    m.train(); ref.train()
    y = m(a, r); y.square().sum().backward()
    y2 = ref(a2 + r2); y2.square().sum().backward()
    torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(a.grad, a2.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r.grad, r2.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(m.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(m.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-3)
    # dropout draws a different mask on every call, and the same one for the same generator state
    m.p = 0.5
    C = native()
    w, b = m.weight.detach(), m.bias.detach()
    torch.manual_seed(5); m1 = C.ln_forward(a.detach(), r.detach(), w, b, 0.5, True, 1e-5)[4]
    m2 = C.ln_forward(a.detach(), r.detach(), w, b, 0.5, True, 1e-5)[4]
    torch.manual_seed(5); m3 = C.ln_forward(a.detach(), r.detach(), w, b, 0.5, True, 1e-5)[4]
    assert not torch.equal(m1, m2) and torch.equal(m1, m3)
    # eval mode: no dropout, no mask
    assert C.ln_forward(a.detach(), r.detach(), w, b, 0.5, False, 1e-5)[4].numel() == 0


@pytest.mark.gpu
def test_fresh_masks_under_cuda_graph_replay():
    dev = torch.device("cuda:0")
    C = native()
    a = torch.randn(64, 512, device=dev); r = torch.randn(64, 512, device=dev)
    w = torch.ones(512, device=dev); b = torch.zeros(512, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        C.ln_forward(a, r, w, b, 0.5, True, 1e-5)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = C.ln_forward(a, r, w, b, 0.5, True, 1e-5)
    g.replay(); torch.cuda.synchronize(); m1 = out[4].clone()
    g.replay(); torch.cuda.synchronize(); m2 = out[4].clone()
    assert not torch.equal(m1, m2)
    assert abs(m1.float().mean().item() - 0.5) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_branch_bias_and_its_gradient(dtype, p):
    """layer_norm(res + dropout(a + b)): bias added in the kernel, d b = column sums of d a from the backward kernel."""
    C = native()
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    shape = (333, 768)
    a = torch.randn(shape, device=dev).to(dtype)
    r = torch.randn(shape, device=dev).to(dtype)
    bb = torch.randn(shape[-1], device=dev).to(dtype)
    w = (1.0 + 0.1 * torch.randn(shape[-1], device=dev)).to(dtype)
    b = (0.1 * torch.randn(shape[-1], device=dev)).to(dtype)
    y, s, mean, rstd, mask = C.ln_forward(a, r, w, b, p, True, 1e-5, bb)
    keep = mask.bool() if p > 0 else None
    a_plus = (a.float() + bb.float()).to(dtype)                       # the kernel rounds a + b to the storage dtype
    y_ref, s_ref = _reference(a_plus, r, w, b, keep, p, 1e-5)
    stol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=8e-3, atol=1e-2)
    torch.testing.assert_close(s.float(), s_ref, **stol)
    ytol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    torch.testing.assert_close(y.float(), F.layer_norm(s.float(), (shape[-1],), w.float(), b.float(), 1e-5), **ytol)
    dy = torch.randn(shape, device=dev).to(dtype)
    d_res, d_a, dgamma, dbeta, dbias = C.ln_backward(dy, s, mean, rstd, w, mask, p, True)
    ref = d_a.float().sum(0)
    tol = dict(rtol=1e-4, atol=1e-3) if dtype == torch.float32 else dict(rtol=2e-2, atol=0.3)
    torch.testing.assert_close(dbias.float(), ref, **tol)
    assert C.ln_backward(dy, s, mean, rstd, w, mask, p, False)[4].numel() == 0
    # through autograd
    a2, bb2 = a.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    torch.manual_seed(9)
    out = dropout_add_layer_norm(a2, r, w, b, p, True, 1e-5, branch_bias=bb2)
    out.backward(dy)
    torch.testing.assert_close(bb2.grad.float(), a2.grad.float().sum(0), **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2048, 4096), (5, 7, 72), (1, 8), (130, 520)])
def test_bias_gelu_forward_backward(dtype, shape):
    from dear_pytorch_b200.ops.bias_gelu import bias_gelu
    dev = torch.device("cuda:0")
    torch.manual_seed(6)
    z = torch.randn(shape, device=dev).to(dtype).requires_grad_(True)
    b = torch.randn(shape[-1], device=dev).to(dtype).requires_grad_(True)
    dh = torch.randn(shape, device=dev).to(dtype)
    h = bias_gelu(z, b)
    h.backward(dh)
    z32, b32 = z.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    zb = (z32 + b32)
    zb = zb + (zb.to(dtype).float() - zb).detach()                    # the kernel rounds z + b to the storage dtype
    h_ref = F.gelu(zb)
    h_ref.backward(dh.float())
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    torch.testing.assert_close(h.float(), h_ref, **tol)
    torch.testing.assert_close(z.grad.float(), z32.grad, **tol)
    rows = z.numel() // shape[-1]
    btol = dict(rtol=1e-4, atol=1e-3) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2 * max(1.0, rows ** 0.5))
    torch.testing.assert_close(b.grad.float(), b32.grad, **btol)
