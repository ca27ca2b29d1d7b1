"""Kernel-level numerics: BucketSet.reduce_scatter / allgather_update against a plain PyTorch fp32
reference of the same op — odd segment sizes (vector tails), absent gradients (zero fill), in-place
segments, several hyper-parameter segments, momentum / nesterov / weight decay, fp32 and bf16.

The same test body runs on the host emulation (CPU, always) and on the CUDA kernels (gpu marker)."""
import pytest
import torch

from _mp import run_ranks


def reference_update(p, g, buf, first, lr, wd, mom, damp, nest):
    g = g + wd * p if wd else g.clone()
    if mom > 0:
        buf = g.clone() if first else mom * buf + (1 - damp) * g
        g = g + mom * buf if nest else buf
    return p - lr * g, buf


def kernel_worker(rank, world, use_cuda, dtype_name, seed, grad_tol=None):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    comm = dear.communicator()
    dev = dear.device()
    tdt = torch.float32 if dtype_name == "fp32" else torch.bfloat16
    es = 4 if dtype_name == "fp32" else 2
    g = torch.Generator().manual_seed(seed)
    # parameters of odd sizes; every start is 256-byte aligned like the planner does
    numels = [37, 1000, 4099, 3, 70001]
    align = 256 // es
    starts, off = [], 0
    for n in numels:
        off = (off + align - 1) // align * align
        starts.append(off)
        off += n
    quantum = world * (128 // es)
    padded = (off + quantum - 1) // quantum * quantum
    shard = padded // world
    bs = C.BucketSet(comm, [padded], C.DT_F32 if dtype_name == "fp32" else C.DT_BF16, True)
    pbuf, gbuf = bs.param_buffer(0), bs.grad_buffer(0)
    full_p = torch.zeros(padded)
    for s, n in zip(starts, numels):
        full_p[s:s + n] = torch.randn(n, generator=g)
    pbuf.copy_(full_p.to(tdt))
    full_p = pbuf.float().cpu().clone()                          # what the kernel sees after rounding
    gs = torch.zeros(shard, device=dev)
    mom = torch.zeros(shard, device=dev)
    master = pbuf[rank * shard:(rank + 1) * shard].float().clone() if dtype_name != "fp32" else None
    bs.set_shards(0, gs, mom, master)
    # hyper segments: params 0-1 group A, 2-3 group B (nesterov), 4 group C (no momentum)
    hyp = [(starts[2], 0.1, 0.01, 0.9, 0.0, 0), (starts[4], 0.05, 0.0, 0.8, 0.0, 1), (padded, 0.2, 0.001, 0.0, 0.0, 0)]
    bs.set_hyper(0, [h[0] for h in hyp], [h[1] for h in hyp], [h[2] for h in hyp], [h[3] for h in hyp],
                 [h[4] for h in hyp], [h[5] for h in hyp])
    ref_p = full_p.clone()
    ref_buf = torch.zeros(padded)
    results = []
    for step in range(3):
        # this rank's gradients: param 3 is absent on step 1 (zero fill), param 1 is "in place" on step 2
        gen = torch.Generator().manual_seed(1000 * step + 7)
        all_rank_grads = [[torch.randn(n, generator=gen).to(tdt) for n in numels] for _ in range(world)]
        mine = [t.to(dev) for t in all_rank_grads[rank]]
        src, flags = [], []
        for i, t in enumerate(mine):
            if step == 1 and i == 3:
                src.append(0); flags.append(C.SEG_ZERO_FILL)
            elif step == 2 and i == 1:
                gbuf[starts[i]:starts[i] + numels[i]].copy_(t)
                src.append(0); flags.append(0)
            else:
                src.append(t.data_ptr()); flags.append(0)
        bs.set_pack(0, src, [s * es for s in starts], [n * es for n in numels], flags)
        bs.reduce_scatter(0, True)
        bs.allgather_update(0, True, step == 0, True, False)
        bs.synchronize()
        comm.check_status()
        # reference
        summed = torch.zeros(padded)
        for r in range(world):
            for i, (s, n) in enumerate(zip(starts, numels)):
                if step == 1 and i == 3:
                    continue
                summed[s:s + n] += all_rank_grads[r][i].float()
        avg = summed / world
        start = 0
        for end, lr, wd, m, damp, nest in hyp:
            sl = slice(start, end)
            ref_p[sl], ref_buf[sl] = reference_update(ref_p[sl], avg[sl], ref_buf[sl], step == 0, lr, wd, m, damp, bool(nest))
            start = end
        got_shard_grad = gs.cpu()
        gt = grad_tol or dict(rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(got_shard_grad, avg[rank * shard:(rank + 1) * shard], **gt)
        got = pbuf.float().cpu()
        if dtype_name == "fp32":
            torch.testing.assert_close(got, ref_p, rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(master.cpu(), ref_p[rank * shard:(rank + 1) * shard], **gt)
            pt = dict(rtol=8e-3, atol=1e-6) if grad_tol is None else dict(rtol=2e-2, atol=1e-2)
            torch.testing.assert_close(got, ref_p.to(torch.bfloat16).float(), **pt)   # <= 1 bf16 ulp
        results.append(float(got.abs().sum()))
    return results


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_emulated_kernels_match_reference(world, dtype_name):
    outs = run_ranks(kernel_worker, world=world, backend="emu", args=(False, dtype_name, 5))
    assert all(o == outs[0] for o in outs)


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("world", [2, 3])
def test_emulated_pipelined_reduce_scatter_matches_reference(world, dtype_name):
    """Host emulation of the stripe-pipelined Kernel A (emu.cpp mirrors rs_pipe.cu: stripe-major work list, one
    RS_READY flag value per stripe): several stripes, segment tails, zero-fill and in-place segments."""
    env = {"DEAR_RS_ALGO": "pipe", "DEAR_STRIPE_MB": "0.0625"}
    outs = run_ranks(kernel_worker, world=world, backend="emu", args=(False, dtype_name, 5), extra_env=env)
    assert all(o == outs[0] for o in outs)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("world", [1, 2])
def test_cuda_kernels_match_reference(world, dtype_name):
    outs = run_ranks(kernel_worker, world=world, backend="b200", args=(True, dtype_name, 5),
                     extra_env={"DEAR_SPIN_TIMEOUT_S": "15"}, timeout=300)
    assert all(o == outs[0] for o in outs)


PIPE_ENV = {"DEAR_SPIN_TIMEOUT_S": "15", "DEAR_RS_ALGO": "pipe", "DEAR_PIPE_MIN_MB": "0", "DEAR_STRIPE_MB": "0.0625"}


@pytest.mark.gpu
@pytest.mark.parametrize("world,dtype_name", [(2, "fp32"), (2, "bf16"), (4, "fp32")])
def test_cuda_pipelined_reduce_scatter_matches_reference(world, dtype_name):
    """Kernel A, stripe-pipelined variant (csrc/rs_pipe.cu: pack warps + TMA bulk-copy pull ring + shared-memory
    reduce) forced onto the small test bucket with 64 KB stripes, so several stripes, partial chunks, segment tails,
    zero-fill and in-place segments all go through it; same fp32 oracle as the one-shot kernel."""
    ngpu = torch.cuda.device_count()
    if world > 2 and ngpu < world:
        pytest.skip("the %d-rank case needs %d GPUs (2 ranks may share one)" % (world, world))
    outs = run_ranks(kernel_worker, world=world, backend="b200", args=(True, dtype_name, 5), extra_env=PIPE_ENV, timeout=300)
    assert all(o == outs[0] for o in outs)


def plan_worker(rank, world):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    bs = C.BucketSet(dear.communicator(), [world * 65536, world * 16 * 1024 * 1024, world * 32 * 1024 * 1024], C.DT_F32, True)
    return [bs.rs_plan(0), bs.rs_plan(1), bs.rs_plan(2)]


@pytest.mark.gpu
def test_reduce_scatter_algorithm_is_picked_per_bucket_size():
    env = {"DEAR_SPIN_TIMEOUT_S": "15", "DEAR_PIPE_MIN_MB": "200"}
    small, mid, big = run_ranks(plan_worker, world=2, backend="b200", extra_env=env, timeout=300)[0]
    assert small.startswith("oneshot:grid=8:"), small       # 512 KB: latency-bound, few CTAs
    assert mid.startswith("oneshot:grid=128:"), mid         # 128 MB: the pack phase wants the wide grid
    assert big.startswith("pipe") and "stripes=16" in big, big      # 256 MB >= DEAR_PIPE_MIN_MB: stripe-pipelined TMA pull


def nvls_worker(rank, world, dtype_name):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    probe = C.BucketSet(dear.communicator(), [world * 4096], C.DT_F32, True)
    if not probe.has_multicast():
        return "no-multicast"
    del probe
    # multimem.ld_reduce on a 16-bit bucket accumulates in fp32 inside the switch but RETURNS the element type:
    # the reduced gradient carries one bf16 rounding (2^-8 relative) that the P2P kernels do not have
    tol = dict(rtol=1e-2, atol=1e-2) if dtype_name == "bf16" else None
    return kernel_worker(rank, world, True, dtype_name, 5, tol)


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
def test_cuda_nvls_multicast_kernels_match_reference(dtype_name):
    """The MC=true instantiations: multimem.ld_reduce in Kernel A, multimem.st in Kernel B (VMM arena bound to an
    NVLS multicast object).  Needs >= 2 GPUs behind an NVSwitch."""
    env = {"DEAR_SPIN_TIMEOUT_S": "15", "DEAR_PROVIDER": "vmm", "DEAR_MULTICAST": "1", "DEAR_RS_ALGO": "nvls"}
    outs = run_ranks(nvls_worker, world=2, backend="b200", args=(dtype_name,), extra_env=env, timeout=300)
    if outs[0] == "no-multicast":
        pytest.skip("this box cannot create an NVLS multicast object")
    assert all(o == outs[0] for o in outs)


def adam_worker(rank, world, use_cuda, dtype_name, seed):
    """Kernel B with the Adam / AdamW epilogue against torch.optim.Adam(W) on the averaged gradient."""
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    comm = dear.communicator()
    dev = dear.device()
    tdt = torch.float32 if dtype_name == "fp32" else torch.bfloat16
    es = 4 if dtype_name == "fp32" else 2
    numels = [513, 4099, 70001]
    align = 256 // es
    starts, off = [], 0
    for n in numels:
        off = (off + align - 1) // align * align
        starts.append(off)
        off += n
    quantum = world * (128 // es)
    padded = (off + quantum - 1) // quantum * quantum
    shard = padded // world
    bs = C.BucketSet(comm, [padded], C.DT_F32 if dtype_name == "fp32" else C.DT_BF16, True)
    pbuf = bs.param_buffer(0)
    g = torch.Generator().manual_seed(seed)
    full_p = torch.zeros(padded)
    for s, n in zip(starts, numels):
        full_p[s:s + n] = torch.randn(n, generator=g)
    pbuf.copy_(full_p.to(tdt))
    full_p = pbuf.float().cpu().clone()
    gs = torch.zeros(shard, device=dev)
    m = torch.zeros(shard, device=dev)
    v = torch.zeros(shard, device=dev)
    master = pbuf[rank * shard:(rank + 1) * shard].float().clone() if dtype_name != "fp32" else None
    bs.set_shards(0, gs, m, master, v)
    bs.set_step(0, 0)
    # params 0-1: Adam with L2 weight decay; param 2: AdamW
    hyp = [(starts[2], 1e-2, 1e-2, 0.9, 0.999, 1e-8, C.OPT_ADAM), (padded, 5e-3, 5e-2, 0.8, 0.95, 1e-6, C.OPT_ADAMW)]
    bs.set_hyper(0, [h[0] for h in hyp], [h[1] for h in hyp], [h[2] for h in hyp], [h[3] for h in hyp],
                 [0.0] * len(hyp), [0] * len(hyp), opt=[h[6] for h in hyp], beta2=[h[4] for h in hyp], eps=[h[5] for h in hyp])
    ref_params = [torch.nn.Parameter(full_p[s:s + n].clone()) for s, n in zip(starts, numels)]
    ref_opts = [torch.optim.Adam(ref_params[:2], lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8),
                torch.optim.AdamW(ref_params[2:], lr=5e-3, weight_decay=5e-2, betas=(0.8, 0.95), eps=1e-6)]
    out = []
    for step in range(4):
        gen = torch.Generator().manual_seed(1000 * step + 11)
        all_rank_grads = [[torch.randn(n, generator=gen).to(tdt) for n in numels] for _ in range(world)]
        mine = [t.to(dev) for t in all_rank_grads[rank]]
        bs.set_pack(0, [t.data_ptr() for t in mine], [s * es for s in starts], [n * es for n in numels], [0] * len(numels))
        bs.reduce_scatter(0, True)
        bs.allgather_update(0, True, step == 0, True, False)
        bs.synchronize()
        comm.check_status()
        for i, p in enumerate(ref_params):
            p.grad = sum(all_rank_grads[r][i].float() for r in range(world)) / world
        for o in ref_opts:
            o.step()
        got = pbuf.float().cpu()
        for i, (s, n) in enumerate(zip(starts, numels)):
            want = ref_params[i].detach()
            if dtype_name == "fp32":
                torch.testing.assert_close(got[s:s + n], want, rtol=2e-5, atol=2e-6)
            else:
                torch.testing.assert_close(got[s:s + n], want.to(torch.bfloat16).float(), rtol=8e-3, atol=1e-5)
        out.append(float(got.abs().sum()))
    return out


@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("world", [1, 3])
def test_emulated_adam_kernel(world, dtype_name):
    outs = run_ranks(adam_worker, world=world, backend="emu", args=(False, dtype_name, 9))
    assert all(o == outs[0] for o in outs)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
@pytest.mark.parametrize("world", [1, 2])
def test_cuda_adam_kernel(world, dtype_name):
    outs = run_ranks(adam_worker, world=world, backend="b200", args=(True, dtype_name, 9),
                     extra_env={"DEAR_SPIN_TIMEOUT_S": "15"}, timeout=300)
    assert all(o == outs[0] for o in outs)


def pieces_worker(rank, world):
    """Host-side work list of the stripe-pipelined Kernel A (BucketSet::set_pack): forced onto the emulation backend."""
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200 import ops
    C = ops.require_native()
    es, padded = 4, world * 160 * 1024           # 640 KB per shard at fp32
    bs = C.BucketSet(dear.communicator(), [padded], C.DT_F32, True)
    gs = torch.zeros(padded // world)
    bs.set_shards(0, gs, None, None)
    # segments with odd sizes, a gap, a zero-fill and an "already in place" one; the last one crosses several stripes
    numels = [1000, 37, 50001, 8, 200000]
    starts, off = [], 0
    for n in numels:
        off = (off + 63) // 64 * 64
        starts.append(off)
        off += n
    assert off <= padded
    srcs = [torch.randn(n) for n in numels]
    ptrs = [t.data_ptr() for t in srcs]
    ptrs[3] = 0                                   # in place: dropped from the list
    flags = [0, C.SEG_ZERO_FILL, 0, 0, 0]
    ptrs[1] = 0
    bs.set_pack(0, ptrs, [s * es for s in starts], [n * es for n in numels], flags)
    return bs.rs_plan(0), bs.pack_pieces(0), [s * es for s in starts], [n * es for n in numels], ptrs


def test_pipelined_work_list_covers_the_pack_table_stripe_major():
    env = {"DEAR_RS_ALGO": "pipe", "DEAR_STRIPE_MB": "0.25"}
    plan, pieces, starts, nbytes, ptrs = run_ranks(pieces_worker, world=2, backend="emu", extra_env=env)[0]
    assert plan.startswith("pipe"), plan
    fields = dict(kv.split("=") for kv in plan.split(":")[1:])
    nstripes, cs = int(fields["stripes"]), int(fields["stripe_bytes"])
    SB = 160 * 1024 * 4
    assert nstripes > 1 and cs % 32768 == 0 and nstripes * cs >= SB
    # stripe-major order, bounded piece size, no piece crosses a stripe or shard boundary
    assert [p[3] for p in pieces] == sorted(p[3] for p in pieces)
    for src, dst, n, k, fl in pieces:
        assert 0 < n <= 32768
        in_shard = dst % SB
        assert in_shard // cs == k and (in_shard + n - 1) // cs == k and (dst + n - 1) // SB == dst // SB
    # exact cover of every listed segment (the in-place one is absent), with matching source offsets
    covered = {}
    for src, dst, n, k, fl in pieces:
        for i, (s0, nb) in enumerate(zip(starts, nbytes)):
            if s0 <= dst < s0 + nb:
                assert dst + n <= s0 + nb
                covered.setdefault(i, []).append((dst, n))
                if fl & 1:
                    assert i == 1
                else:
                    assert src == ptrs[i] + (dst - s0)
    assert sorted(covered) == [0, 1, 2, 4]
    for i, spans in covered.items():
        spans.sort()
        assert spans[0][0] == starts[i] and sum(n for _, n in spans) == nbytes[i]
        assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
