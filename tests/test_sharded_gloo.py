"""CPU suite: the N>1 host logic (slab partitioning, single all-reduce for reductions over the sharded dim) with
world_size 2 over gloo.  The per-rank compute is injected (the C restatement of the reference stands in for the
CUDA engine, which cannot run here); what is under test is strided.jl_b200/sharded.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import sb, col_major_strides
from oracle import ref as oref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(tokens, opc, initop, dims, arrays):
    views = sb.promoteshape(dims, *arrays)
    oref.mapreduce(sb.make_desc(tokens, opc, 0, 0.0, dims, views), 2)


def _np_alloc(like, dtype_code, shape):
    n = int(np.prod(shape)) if shape else 1
    return sb.StridedView(np.zeros(max(n, 1), dtype=sb.view.sb_to_numpy_dtype(dtype_code)), tuple(shape), col_major_strides(shape))


def _np_fill(out, opc):
    out.parent[...] = {1: 0.0, 2: 1.0, 3: np.inf, 4: -np.inf}[opc]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from strided_jl_b200 import sharded
    try:
        g, m = 8, 37
        rng = np.random.default_rng(1234)
        full = rng.standard_normal(g * m * m)
        A = sb.StridedView(full, (g, m, m), (1, g, g * m))
        A3 = full.reshape((g, m, m), order="F")
        kw = dict(compute=_oracle_compute, alloc=_np_alloc, fill=_np_fill)
        # (1) config-5 shape: kept dim 0 sharded, reduce dims (1,2): no collective, each rank owns g/world outputs
        loc = sharded.shard_view(A, 0, rank, world)
        lo, hi = sharded.shard_range(g, rank, world)
        out = sharded.sharded_mapreduce("abs2", "+", loc, dims=(1, 2), shard_dim=0, **kw)
        np.testing.assert_allclose(out.to_numpy().reshape(-1), (A3[lo:hi] ** 2).sum(axis=(1, 2)), rtol=1e-12)
        # (2) reduced dim sharded: one all-reduce of 8 elements gives every rank the full answer
        loc = sharded.shard_view(A, 2, rank, world)
        out = sharded.sharded_mapreduce("abs2", "+", loc, dims=(1, 2), shard_dim=2, **kw)
        np.testing.assert_allclose(out.to_numpy().reshape(-1), (A3 ** 2).sum(axis=(1, 2)), rtol=1e-12)
        # (3) complete reductions: one all-reduce of ONE element
        s = sharded.sharded_mapreduce("identity", "+", loc, shard_dim=2, **kw)
        assert abs(s - full.sum()) < 1e-9
        mx = sharded.sharded_mapreduce("abs", "max", loc, shard_dim=2, **kw)
        assert mx == np.abs(full).max()
        # uneven slabs
        assert [sharded.shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_sharded_reductions_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
