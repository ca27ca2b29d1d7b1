"""The native communicator's operation family on the host-emulation backend, and the Comm facade /
tensor-fusion helpers on both CPU backends (counterpart of common/comm_core/tests/test_comm.py, which
only prints norms for a human to eyeball)."""
import pytest
import torch

from _mp import run_ranks


def ops_worker(rank, world):
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200.parallel.comm import Comm
    c = Comm()
    out = {}
    t = torch.arange(17.0) * (rank + 1)
    c.syncStream(c.allReduce(t, 1.0))
    out["allreduce"] = t.clone()
    t2 = torch.arange(17.0) * (rank + 1)
    c.syncStream(c.allReduceRB(t2, 1.0))
    out["allreduce_rb"] = t2.clone()
    n = 8 * world
    t3 = torch.arange(float(n)) + rank
    c.syncStream(c.allReduceRSAG(t3, 1.0 / world))
    out["rsag_avg"] = t3.clone()
    send = torch.arange(float(n)) * (rank + 1)
    recv = torch.zeros(n // world)
    c.syncStream(c.reduceScatter(send, recv, 1.0))
    out["reduce_scatter"] = recv.clone()
    gathered = torch.zeros(3 * world)
    c.syncStream(c.allGather(torch.full((3,), float(rank)), gathered))
    out["allgather"] = gathered.clone()
    b = torch.arange(5, dtype=torch.int64) * (rank + 2)
    c.syncStream(c.bcast(b, world - 1))
    out["bcast_int64"] = b.clone()
    r = torch.ones(6) * (rank + 1)
    c.syncStream(c.reduce(r, 0, 1.0))
    out["reduce_root0"] = r.clone()
    s, v = torch.full((4,), float(rank)), torch.zeros(4)
    c.syncStream(c.sendrecv(s, v, (rank + 1) % world))
    out["sendrecv"] = v.clone()
    outs = [torch.zeros(2, 2), torch.zeros(600, 600)]
    c.multiBcast([torch.ones(2, 2), torch.ones(600, 600)], outs, lambda t, o: o.copy_(t * 3))
    out["multibcast"] = (float(outs[0].sum()), float(outs[1].sum()))
    out["free_streams"] = c.getNumOfFreeStreams()
    avg = dear.allreduce(torch.tensor([float(rank)]))
    out["api_allreduce"] = float(avg)
    dear.barrier()
    return out


@pytest.mark.parametrize("backend", ["emu", "gloo"])
def test_collective_family(backend):
    world = 2
    outs = run_ranks(ops_worker, world=world, backend=backend)
    s = sum(range(1, world + 1))
    for r, o in enumerate(outs):
        torch.testing.assert_close(o["allreduce"], torch.arange(17.0) * s)
        torch.testing.assert_close(o["allreduce_rb"], torch.arange(17.0) * s)
        torch.testing.assert_close(o["rsag_avg"], torch.arange(16.0) + 0.5)
        per = 16 // world
        torch.testing.assert_close(o["reduce_scatter"], (torch.arange(16.0) * s)[r * per:(r + 1) * per])
        torch.testing.assert_close(o["allgather"], torch.arange(world).repeat_interleave(3).float())
        assert torch.equal(o["bcast_int64"], torch.arange(5) * (world - 1 + 2))
        if r == 0:
            torch.testing.assert_close(o["reduce_root0"], torch.ones(6) * s)
        torch.testing.assert_close(o["sendrecv"], torch.full((4,), float((r + 1) % world)))
        assert o["multibcast"] == (12.0, 3.0 * 600 * 600)
        assert o["free_streams"] >= 1
        assert abs(o["api_allreduce"] - 0.5) < 1e-6


def fusion_worker(rank, world):
    from dear_pytorch_b200.parallel.tensorfusion import (CollectiveOp, CommReduceScatter, MergedCommCollective,
                                                         MergedCommReduce, TensorGroup)
    names = ["a", "b", "c", "d", "e"]
    sizes = {"a": 3, "b": 5, "c": 2, "d": 7, "e": 1}
    tg = TensorGroup(names, num_nearby_layers=2, sizes=sizes)
    assert tg.groups == [["a", "b"], ["c", "d"], ["e"]]
    assert tg.push_tensor("a", torch.ones(3))[1] is None
    gname, buf = tg.push_tensor("b", torch.full((5,), 2.0))
    assert gname == "group-0" and buf.tolist() == [1.0] * 3 + [2.0] * 5
    tg.regroup_by_flags([1, 0, 0, 1, 0])
    assert tg.groups == [["a", "b", "c"], ["d", "e"]]

    tensors = {n: torch.full((sizes[n],), float(rank + 1)) for n in names}
    mc = MergedCommCollective(names, merge=True, op=CollectiveOp.ALL_REDUCE, num_nearby_layers=2)
    mc.init_tensor_group(names, sizes)
    launched = [mc.collective_async_(n, tensors[n]) for n in names]
    assert sum(h is not None for h in launched) == 3
    res = mc.synchronize()
    total = float(sum(range(1, world + 1)))
    assert all(torch.equal(res[n], torch.full((sizes[n],), total)) for n in names)

    sym = torch.tensor([[1.0, 2.0], [2.0, 3.0]]) * (rank + 1)
    ms = MergedCommCollective(merge=False, symmetric=True, op=CollectiveOp.ALL_REDUCE)
    ms.collective_async_("k", sym)
    ms.synchronize()
    assert torch.equal(sym, torch.tensor([[1.0, 2.0], [2.0, 3.0]]) * total)

    red = MergedCommReduce(merge=False, op=CollectiveOp.REDUCE)
    x = torch.ones(4) * (rank + 1)
    red.reduce_async_("x", x, 0)
    red.synchronize()
    if rank == 0:
        assert torch.equal(x, torch.full((4,), total))

    rs = CommReduceScatter(op=CollectiveOp.REDUCE_SCATTER)
    ag = CommReduceScatter(op=CollectiveOp.ALL_GATHER)
    pad = torch.arange(8.0) * (rank + 1)
    shard = torch.zeros(8 // world)
    rs.collective_async_("g", pad, shard)
    rs.synchronize()
    ag.collective_async_("g", pad, shard)
    ag.synchronize()
    assert torch.equal(pad, torch.arange(8.0) * total)         # RS followed by AG == all-reduce
    return True


@pytest.mark.parametrize("backend", ["emu", "gloo"])
def test_tensorfusion_helpers(backend):
    assert all(run_ranks(fusion_worker, world=2, backend=backend))


def gtopk_worker(rank, world):
    from dear_pytorch_b200.parallel.baselines.gtopk import gtopk_sparse_recursive_allreduce
    from dear_pytorch_b200.parallel.comm import Comm
    torch.manual_seed(rank)
    n, k = 64, 6
    dense = torch.randn(n)
    idx = torch.topk(dense.abs(), k)[1]
    vals, gidx = gtopk_sparse_recursive_allreduce(Comm(), dense[idx], idx, n, k)
    return dense, idx, vals, gidx


def test_gtopk_sparse_allreduce_is_rank_consistent():
    outs = run_ranks(gtopk_worker, world=2, backend="emu")
    assert torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][3], outs[1][3])
    # oracle: sum of the two sparsified vectors, then top-k by magnitude
    full = torch.zeros(64)
    for dense, idx, _, _ in outs:
        full[idx] += dense[idx]
    k = 6
    top = torch.topk(full.abs(), k)[1]
    got = torch.zeros(64)
    got[outs[0][3]] = outs[0][2]
    want = torch.zeros(64)
    want[top] = full[top]
    torch.testing.assert_close(got, want)


def timeout_worker(rank, world):
    """Failure detection: a peer that never joins a collective is reported, not waited for forever."""
    import dear_pytorch_b200 as dear
    comm = dear.communicator()
    if rank == 0:
        t = torch.ones(4)
        comm.allReduce(t, 1.0)           # rank 1 never calls it: bounded spin, then the status word is set
        try:
            comm.check_status()
        except RuntimeError as e:
            return "timeout" if "timed out" in str(e) else str(e)
        return "no error"
    import time
    time.sleep(3.0)
    return "skipped"


def test_missing_peer_is_detected_not_hung():
    outs = run_ranks(timeout_worker, world=2, backend="emu", extra_env={"DEAR_SPIN_TIMEOUT_S": "1"}, timeout=60)
    assert outs[0] == "timeout" and outs[1] == "skipped"


def test_backend_selection_single_and_multi_node():
    from dear_pytorch_b200.runtime import select_backend
    assert select_backend(None, False, 2, 2) == "gloo"
    assert select_backend(None, True, 8, 8) == "b200"
    assert select_backend(None, True, 16, 8) == "nccl"          # two nodes: peer memory does not span them
    assert select_backend("nccl", True, 16, 8) == "nccl"
    assert select_backend("emu", False, 3, 3) == "emu"
    with pytest.raises(RuntimeError):
        select_backend("b200", True, 16, 8)
    with pytest.raises(ValueError):
        select_backend("mpi", True, 8, 8)


def bos_worker(rank, world):
    import dear_pytorch_b200 as dear
    import torch.nn as nn
    torch.manual_seed(rank)                       # different initial weights per rank
    model = nn.Linear(5, 3)
    opt = torch.optim.SGD(model.parameters(), lr=0.1 * (rank + 1), momentum=0.9)
    if rank == 0:                                 # only the root has optimizer state (e.g. it resumed a checkpoint)
        model(torch.ones(2, 5)).sum().backward()
        opt.step()
    dear.broadcast_parameters(model.state_dict(), 0)
    dear.broadcast_optimizer_state(opt, 0)
    bufs = [opt.state[p]["momentum_buffer"].clone() for p in model.parameters()]
    return opt.param_groups[0]["lr"], bufs, [p.detach().clone() for p in model.parameters()]


def test_broadcast_optimizer_state_to_a_rank_without_state():
    outs = run_ranks(bos_worker, world=2, backend="gloo")
    assert outs[0][0] == outs[1][0] == 0.1
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b) and a.abs().sum() > 0
    for a, b in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, b)


def streams_worker(rank, world):
    """Comm(nstreams=k) grows the shared native communicator (reference _extendComms, communicator.cpp:85-95)."""
    import torch
    import dear_pytorch_b200 as dear
    from dear_pytorch_b200.parallel.comm import Comm
    native = dear.communicator()
    before = native.numStreams()
    c = Comm(nstreams=3)
    assert native.numStreams() == max(before, 3)
    # three operations in flight on three different slots (each has its own staging arena and flags)
    ts = [torch.full((1000,), float(rank + 1 + i)) for i in range(3)]
    hs = [c.allReduce(t, 1.0) for t in ts]
    assert len(set(hs)) == 3
    for h in hs:
        c.syncStream(h)
    return [float(t[0]) for t in ts]


def test_comm_nstreams_grows_the_native_communicator():
    outs = run_ranks(streams_worker, world=2, backend="emu")
    assert outs[0] == outs[1] == [3.0, 5.0, 7.0]


def comm_core_worker(rank, world):
    """The flows of the reference's smoke script (common/comm_core/tests/test_comm.py:11-65) through the drop-in
    ``comm_core`` module — with assertions instead of prints."""
    import comm_core
    comm_core.init()
    assert (comm_core.rank(), comm_core.size()) == (rank, world)
    dev = __import__("dear_pytorch_b200").device()
    comm = comm_core.Communicator(1)
    total = float(sum(range(1, world + 1)))
    # allreduce()
    t = torch.full((2, 2), float(rank + 1), device=dev)
    comm.allReduce(t)
    comm.synchronize()
    assert torch.equal(t.cpu(), torch.full((2, 2), total))
    # reducescatter(): RS -> AG reproduces the all-reduce
    send = torch.arange(16.0, device=dev) * (rank + 1)
    results = torch.zeros_like(send)
    recv = send.new_zeros(send.numel() // world)
    comm.reduceScatter(send, recv)
    comm.allGather(recv, results)
    comm.allReduce(send)
    comm.synchronize()
    assert float((results - send).norm()) == 0.0
    # decoupleallreduce(): an odd size through the reduce + broadcast composition
    a = torch.arange(17.0, device=dev) * (rank + 1)
    b = a.clone()
    comm.allReduce(a)
    comm.allReduceRB(b)
    comm.synchronize()
    assert float((a - b).norm()) == 0.0
    # bcast()
    x = torch.full((2,), 10.0 if rank == 0 else -1.0, device=dev)
    comm.bcast(x, 0)
    comm.synchronize()
    assert x.tolist() == [10.0, 10.0]
    comm_core.barriar()
    comm.destroy()
    comm.reload()
    comm.allReduce(x)
    comm.synchronize()
    assert x.tolist() == [10.0 * world] * 2
    return True


@pytest.mark.parametrize("backend", ["emu", "gloo"])
def test_comm_core_drop_in_module(backend):
    assert all(run_ranks(comm_core_worker, world=2, backend=backend))
