"""Multi-GPU use of the hot path: one process per GPU (`torch.distributed`, NCCL over NVLink), shards of a large
OUTER dimension per rank, no data-path collective except for reductions over the sharded dimension.

The reference parallelises by bisecting the index space over shared-memory tasks (src/mapreduce.jl:195-227) and,
for complete reductions, by per-task partials + a serial fold (:153-170).  Across GPUs the same two shapes are:

  * map / broadcast / permute, and partial reductions whose KEPT dim is sharded: output elements are disjoint per
    rank -> every rank runs the single-GPU engine on its slab, nothing is exchanged (SURVEY.md section 8 e);
  * reductions over the sharded dim (incl. complete reductions): every rank reduces its slab to a partial of the
    full output shape, then ONE all-reduce of `#outputs` elements (1 element for `sum(A)`) combines them --
    the GPU analog of the per-task slots.  Message sizes are bytes..KB, i.e. latency-bound.

Placement matters more than the collective: a rank's slab must be DENSE in its own HBM (e.g. config 5: GPU g holds
A[g,:,:] as a contiguous 4096x4096 array), otherwise a strided slab of a global array wastes sector bandwidth.
"""
from __future__ import annotations

from . import abi
from .broadcast import trace, result_dtype
from .engine import similar_parent
from .mapreduce import _mapreducedim_, _op_code, _fill_scalar
from .view import StridedView, maybestrided

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(n, rank, world):
    """Contiguous slab [lo, hi) of an extent n for `rank` of `world` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_view(A: StridedView, dim: int, rank: int, world: int) -> StridedView:
    """The slab of a (replicated or host-side) view along `dim` owned by `rank` -- metadata only."""
    lo, hi = shard_range(A.size[dim], rank, world)
    idx = tuple(slice(lo, hi) if d == dim else slice(None) for d in range(A.ndim))
    return A[idx]


def attach_peer_group(group=None, device=None):
    """Map every rank's exchange buffer into this process (CUDA IPC; handles travel through torch.distributed) so that
    reductions over the sharded dim are combined by the library's own kernel over NVLink peer memory instead of an
    NCCL all-reduce (include/strided_b200.h: sb_peer_export / sb_peer_attach / sb_mapreduce_allreduce).  Collective."""
    from .engine import get_engine
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world > abi.SB_PEER_MAX_WORLD:
        raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"peer groups span one NVSwitch domain (<= {abi.SB_PEER_MAX_WORLD} GPUs)")
    eng = get_engine(torch.cuda.current_device() if device is None else device)
    handles = [None] * world
    dist.all_gather_object(handles, eng.peer_export(), group=group)
    eng.peer_attach(rank, world, handles)
    dist.barrier(group)  # nobody pushes into a buffer that is still being reset
    return eng


_REDOP = {abi.SB_OP_ADD: "SUM", abi.SB_OP_MUL: "PRODUCT", abi.SB_OP_MIN: "MIN", abi.SB_OP_MAX: "MAX"}


def _neutral_fill(out, opc):
    if opc == abi.SB_OP_ADD:
        _fill_scalar(out, 0)
    elif opc == abi.SB_OP_MUL:
        _fill_scalar(out, 1)
    elif opc == abi.SB_OP_MIN:
        _fill_scalar(out, float("inf"))
    else:
        _fill_scalar(out, float("-inf"))


def sharded_mapreduce(f, op, A_local, dims=None, shard_dim=None, group=None, compute=None, alloc=None, fill=None, fused=None):
    """mapreduce(f, op, A; dims) where `A_local` is this rank's slab of A along `shard_dim`.

    dims=None  : complete reduction -> Python scalar on every rank (one all-reduce of ONE element).
    dims given : `shard_dim in dims`  -> partial of the full output per rank + one all-reduce of #outputs elements;
                 otherwise           -> the rank's own slab of the output, no collective.
    `compute`, `alloc`, `fill` are injection points for the CPU (gloo) tests; the product path uses the CUDA engine.
    `fused`: combine the ranks' partials with the library's peer-memory kernel (needs `attach_peer_group`) instead of
    an NCCL all-reduce; default: whenever the engine has a peer group and the output is small enough.
    """
    A_local = maybestrided(A_local)
    compute = compute or _mapreducedim_
    alloc = alloc or similar_parent
    fill = fill or _neutral_fill
    tokens = trace(f, 1)
    opc = _op_code(op)
    n = A_local.ndim
    red = tuple(range(n)) if dims is None else ((dims,) if isinstance(dims, int) else tuple(int(d) for d in dims))
    outsize = tuple(1 if d in red else s for d, s in enumerate(A_local.size))
    out = alloc(A_local, result_dtype(tokens, [A_local]), outsize)
    exchange = shard_dim is None or shard_dim in red
    world = dist.get_world_size(group) if (dist is not None and dist.is_initialized()) else 1
    nout = 1
    for s_ in outsize:
        nout *= s_
    if fused is None:
        fused = False
        if compute is _mapreducedim_ and exchange and world > 1 and A_local.is_device and nout <= abi.SB_PEER_MAX_OUT:
            from .engine import get_engine
            fused = get_engine(torch.cuda.current_device()).peer_world == world
    if fused and exchange and world > 1:
        # ONE call: local reduction + exchange of the partials through peer memory + fold in rank order, all ranks;
        # initop = x -> neutral element, so the (uninitialised) output needs no separate fill
        from .broadcast import promoteshape
        from .engine import run_mapreduce
        neutral = {abi.SB_OP_ADD: 0.0, abi.SB_OP_MUL: 1.0, abi.SB_OP_MIN: float("inf"), abi.SB_OP_MAX: float("-inf")}[opc]
        views = promoteshape(A_local.size, out, A_local)
        run_mapreduce(tokens, opc, abi.SB_INIT_CONST, neutral, A_local.size, views, allreduce=True)
        if dims is None:
            return out.to_numpy().reshape(-1)[0].item()
        return out
    fill(out, opc)  # neutral element: every rank contributes op-partials only
    if len(A_local) > 0:
        compute(tokens, opc, None, A_local.size, (out, A_local))
    if exchange and world > 1:
        t = out.parent if torch is not None and isinstance(out.parent, torch.Tensor) else torch.from_numpy(out.parent)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, _REDOP[opc]), group=group)  # the single collective of the path
    if dims is None:
        return out.to_numpy().reshape(-1)[0].item()
    return out


def sharded_map_(f, dst_local, *srcs_local):
    """map!/broadcast/permute on slabs: outputs are disjoint per rank, so this is the single-GPU call (no collective)."""
    from .mapreduce import map_
    return map_(f, dst_local, *srcs_local)
