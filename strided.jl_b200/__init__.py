"""strided.jl_b200 -- B200-native strided map / permute / reduce engine behind Strided.jl's API surface.

Only the hot path of Jutho/Strided.jl is here: the host-side mirror of the reference
interface (StridedView, broadcast capture, map!/mapreduce adapters) over the C ABI of
include/strided_b200.h, whose kernels live in csrc/ (hand-written CUDA for sm_100a).

Import it as ``import strided_jl_b200`` (the root-level alias module; a directory name with a dot cannot be
imported directly).
"""
from . import abi
from .abi import DimensionMismatch, UnsupportedError, NoDeviceError, StridedB200Error, plan_describe
from .view import StridedView, sreshape, sview, isstrided, maybestrided
from .broadcast import (Broadcasted, Ref, Arg, materialize, materialize_, capturestridedargs, promoteshape,
                        make_program, trace, identity, neg, conj, abs_, abs2, real, imag, sqrt, exp, log, sin, cos,
                        tanh, inv, add, sub, mul, div, maximum2, minimum2, lt)
from .engine import get_engine, make_desc, run_mapreduce, run_batch, similar_parent
from .mapreduce import (map_, map, copy_, conj_, adjoint_, transpose_, permutedims_, mapreduce, mapreducedim_,
                        _mapreducedim_, sum, prod, maximum, minimum, rmul_, lmul_, mul_, axpy_, axpby_, mul_generic_, _mul_generic_call)

from . import sharded
from .sharded import sharded_mapreduce, sharded_map_, shard_view, shard_range

__all__ = [n for n in dir() if not n.startswith("__")]
