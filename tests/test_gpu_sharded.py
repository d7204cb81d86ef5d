"""GPU suite: the sharded (one process per GPU) use of the path with the CUDA engine as the per-rank compute.
With >= 2 GPUs the ranks are real NCCL ranks; with one GPU the same code runs at world size 1 (no collective)."""
import os
import socket

import numpy as np
import pytest

from helpers import sb

pytestmark = pytest.mark.gpu


def _check(rank, world, device):
    import torch
    from strided_jl_b200 import sharded
    g, m = 8, 256
    rng = np.random.default_rng(1234)
    full = rng.standard_normal(g * m * m)
    A3 = full.reshape((g, m, m), order="F")
    # config-5 placement: every rank holds its slab DENSE in its own HBM
    lo, hi = sharded.shard_range(g, rank, world)
    slab = torch.from_numpy(np.ascontiguousarray(A3[lo:hi].reshape(-1, order="F"))).to(device)
    loc = sb.StridedView(slab, (hi - lo, m, m), (1, hi - lo, (hi - lo) * m))
    out = sharded.sharded_mapreduce("abs2", "+", loc, dims=(1, 2), shard_dim=0)
    np.testing.assert_allclose(out.to_numpy().reshape(-1), (A3[lo:hi] ** 2).sum(axis=(1, 2)), rtol=1e-12)
    # reduced dim sharded -> one all-reduce of g elements
    lo, hi = sharded.shard_range(m, rank, world)
    slab2 = torch.from_numpy(np.ascontiguousarray(A3[:, :, lo:hi].reshape(-1, order="F"))).to(device)
    loc2 = sb.StridedView(slab2, (g, m, hi - lo), (1, g, g * m))
    out = sharded.sharded_mapreduce("abs2", "+", loc2, dims=(1, 2), shard_dim=2)
    np.testing.assert_allclose(out.to_numpy().reshape(-1), (A3 ** 2).sum(axis=(1, 2)), rtol=1e-12)
    # complete reduction -> one all-reduce of ONE element
    s = sharded.sharded_mapreduce("identity", "+", loc2, shard_dim=2)
    assert abs(s - full.sum()) < 1e-8
    assert sharded.sharded_mapreduce("abs", "max", loc2, shard_dim=2) == np.abs(full).max()
    if world > 1:
        _check_fused(rank, world, device, loc2, A3, full)


def _check_fused(rank, world, device, loc2, A3, full):
    """the same reductions with the partials combined by the library's peer-memory kernel (no NCCL on the path)"""
    import torch
    import torch.distributed as dist
    from strided_jl_b200 import sharded
    eng = sharded.attach_peer_group()
    assert eng.peer_world == world
    launches0 = eng.stats()["launches"]
    out = sharded.sharded_mapreduce("abs2", "+", loc2, dims=(1, 2), shard_dim=2)           # 8 outputs, fused by default now
    np.testing.assert_allclose(out.to_numpy().reshape(-1), (A3 ** 2).sum(axis=(1, 2)), rtol=1e-12)
    assert eng.stats()["launches"] - launches0 == 1, "ONE launch: the exchange is fused into the reduction kernel"
    unfused = sharded.sharded_mapreduce("abs2", "+", loc2, dims=(1, 2), shard_dim=2, fused=False)
    np.testing.assert_allclose(out.to_numpy(), unfused.to_numpy(), rtol=1e-13)
    s = sharded.sharded_mapreduce("identity", "+", loc2, shard_dim=2)
    assert abs(s - full.sum()) < 1e-8
    assert sharded.sharded_mapreduce("abs", "max", loc2, shard_dim=2) == np.abs(full).max()
    assert sharded.sharded_mapreduce("identity", "min", loc2, shard_dim=2) == full.min()
    # every rank folds the same partials in the same (rank) order: bit-identical results everywhere
    mine = torch.tensor([s], dtype=torch.float64, device=device)
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert all(torch.equal(b, both[0]) for b in both)
    # many calls back to back without host synchronisation (epochs, slot parity), accumulating into the same output
    acc = torch.zeros(1, dtype=torch.float64, device=device)
    O = sb.StridedView(acc, loc2.size, (0, 0, 0))
    prog = [(0, 0, 0.0, 0.0), (2, sb.abi.FN["abs2"], 0.0, 0.0)]
    eng.set_sync(False)
    for _ in range(25):
        sb.run_mapreduce(prog, 1, 0, 0.0, loc2.size, [O, loc2], allreduce=True)  # acc = acc + sum over ALL ranks
    torch.cuda.synchronize()
    eng.set_sync(True)
    np.testing.assert_allclose(acc.item(), 25 * (full ** 2).sum(), rtol=1e-12)
    # a collective call captured in a CUDA graph and REPLAYED with different data: the call number lives in device
    # memory and is advanced by the kernel, so every replay gets fresh epochs (a host-side counter baked into the kernel
    # parameters would let a rank accept the previous replay's slots)
    x = torch.zeros(4096, dtype=torch.float64, device=device)
    tot = torch.zeros(1, dtype=torch.float64, device=device)
    X, T = sb.StridedView(x, (4096,), (1,)), sb.StridedView(tot, (4096,), (0,))
    ident = [(0, 0, 0.0, 0.0)]
    eng.set_sync(False)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng.set_stream(side.cuda_stream)
        sb.run_mapreduce(ident, 1, 1, 0.0, (4096,), [T, X], engine=eng, allreduce=True)  # plan + buffers before capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(3):  # three collective calls per replay: both slot parities are exercised
                sb.run_mapreduce(ident, 1, 1, 0.0, (4096,), [T, X], engine=eng, allreduce=True)
        for it in range(6):
            x.fill_(float(it + 1) * (rank + 1))
            if it % 2 == rank % 2:
                torch.cuda._sleep(20_000_000)  # skew the ranks (~10 ms): the fast one must wait, not take stale data
            graph.replay()
            torch.cuda.synchronize()
            want = 4096.0 * (it + 1) * sum(r + 1 for r in range(world))
            assert tot.item() == want, (it, tot.item(), want)
        del graph
    eng.set_sync(True)
    # mixed paths: rank 0 takes the two-kernel path (local reduction + stand-alone exchange kernel), the others the
    # fused one -- same wire format, same results
    if rank == 0:
        os.environ["SB_NO_FUSED_PEER"] = "1"
        peers = eng.peer_world
        eng.reload_env()  # (the SB_* knobs are read at ctx creation / on reload; the peer group survives)
        assert eng.peer_world == peers
    launches0 = eng.stats()["launches"]
    out2 = sharded.sharded_mapreduce("abs2", "+", loc2, dims=(1, 2), shard_dim=2)
    assert eng.stats()["launches"] - launches0 == (2 if rank == 0 else 1)
    assert np.array_equal(out2.to_numpy(), out.to_numpy())
    s2 = sharded.sharded_mapreduce("identity", "+", loc2, shard_dim=2)
    assert s2 == s
    os.environ.pop("SB_NO_FUSED_PEER", None)
    eng.reload_env()
    eng.peer_detach()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _check(rank, world, torch.device("cuda", rank))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_sharded_reductions_on_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        _check(0, 1, torch.device("cuda", 0))
        return
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
