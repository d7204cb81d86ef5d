# StridedB200.jl -- the reference-side binding of libstrided_b200.so.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image nor on the GPU box (SURVEY.md
# section 8c).  It mirrors tests/helpers.py + strided.jl_b200/engine.py one to one, so every call below is
# exercised -- through ctypes instead of ccall -- by the GPU test-suite.
#
# What it does: adds methods for StridedViews whose parent lives in B200 HBM (`B200Vector <: DenseVector`) at the
# single funnel of the hot path, `Strided._mapreduce_fuse!` (reference src/mapreduce.jl:98), so that every
# `@strided` block, `map!`, `permutedims!`, broadcast `copyto!` and `mapreduce` keeps working unchanged and
# lands in `sb_mapreduce`.  Anything the device path declines (SB_E_UNSUPPORTED: opaque closures, exotic
# eltypes) falls back to the original CPU method on a host copy.
module StridedB200

using Strided, StridedViews
import Strided: _mapreduce_fuse!, CaptureArgs, Arg

const LIB = get(ENV, "STRIDED_B200_LIB", "libstrided_b200.so")
const SB_MAX_DIMS, SB_MAX_OPS, SB_MAX_TOKENS = 8, 8, 48

# ---- mirror of include/strided_b200.h -------------------------------------------------------------------
struct SbTok
    kind::Int32; a::Int32; re::Float64; im::Float64
end
struct SbDesc
    ndim::Int32; nops::Int32
    dims::NTuple{SB_MAX_DIMS,Int64}
    strides::NTuple{SB_MAX_OPS,NTuple{SB_MAX_DIMS,Int64}}
    base::NTuple{SB_MAX_OPS,Ptr{Cvoid}}
    dtype::NTuple{SB_MAX_OPS,Int32}
    conj::NTuple{SB_MAX_OPS,Int32}
    ntok::Int32
    prog::NTuple{SB_MAX_TOKENS,SbTok}
    op::Int32; initop::Int32
    init_re::Float64; init_im::Float64
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        rc = ccall((:sb_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), 0, C_NULL, CTX)
        rc == 0 || error("sb_ctx_create failed ($rc): no B200 available")
    end
    return CTX[]
end

# ---- device-resident dense parent: what makes StridedView "device agnostic" (README.md:51-54) ------------
mutable struct B200Vector{T} <: DenseVector{T}
    ptr::Ptr{T}; len::Int
    function B200Vector{T}(::UndefInitializer, n::Integer) where {T}
        p = Ref{Ptr{Cvoid}}()
        rc = ccall((:sb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}), ctx(), n * sizeof(T), p)
        rc == 0 || throw(OutOfMemoryError())
        v = new{T}(Ptr{T}(p[]), n)
        finalizer(x -> ccall((:sb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.ptr), v)
    end
end
Base.size(v::B200Vector) = (v.len,)
Base.pointer(v::B200Vector{T}, i::Integer=1) where {T} = v.ptr + (i - 1) * sizeof(T)
Base.similar(v::B200Vector, ::Type{T}, dims::Dims) where {T} = reshape(B200Vector{T}(undef, prod(dims)), dims)
todevice(a::Array{T}) where {T} = (v = B200Vector{T}(undef, length(a));
    ccall((:sb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), v.ptr, a, sizeof(a)); v)
tohost(v::B200Vector{T}) where {T} = (a = Vector{T}(undef, v.len);
    ccall((:sb_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), a, v.ptr, sizeof(a)); a)

const DeviceView{T,N} = StridedView{T,N,<:B200Vector}

# ---- (f, op, initop) -> program / codes -------------------------------------------------------------------
dtypecode(::Type{Float32}) = 0; dtypecode(::Type{Float64}) = 1
dtypecode(::Type{ComplexF32}) = 2; dtypecode(::Type{ComplexF64}) = 3
dtypecode(::Type) = -1
const FN = Dict{Any,Int32}(identity => 0, (-) => 1, conj => 2, abs => 3, abs2 => 4, real => 5, imag => 6, sqrt => 7,
    exp => 8, log => 9, sin => 10, cos => 11, tanh => 12, inv => 13,
    (+) => 32, (*) => 34, (/) => 35, max => 36, min => 37, (<) => 38)   # binary `-` is 33, resolved by arity below
struct Unsupported <: Exception end

# depth-first, left-to-right: exactly the order in which `consume` (broadcast.jl:86-98) eats the arguments
function program!(toks, c::CaptureArgs, k::Ref{Int})
    nargs = length(c.args)
    f = c.f
    if (f === +) || (f === *)                       # n-ary: left fold, as Julia evaluates it
        program!(toks, c.args[1], k)
        for a in c.args[2:end]
            program!(toks, a, k); push!(toks, SbTok(2, FN[f], 0, 0))
        end
        return
    end
    foreach(a -> program!(toks, a, k), c.args)
    id = (f === -) ? (nargs == 1 ? Int32(1) : Int32(33)) : get(FN, f, Int32(-1))
    id < 0 && throw(Unsupported())
    push!(toks, SbTok(2, id, 0, 0))
end
program!(toks, ::Arg, k::Ref{Int}) = (push!(toks, SbTok(0, k[], 0, 0)); k[] += 1)
program!(toks, x::Union{Integer,Rational,Bool}, k) = push!(toks, SbTok(1, 0, Float64(x), 0))        # weak literal
program!(toks, x::Float32, k) = push!(toks, SbTok(1, 1, x, 0))
program!(toks, x::Float64, k) = push!(toks, SbTok(1, 2, x, 0))
program!(toks, x::Complex, k) = push!(toks, SbTok(1, real(x) isa Float32 ? 1 : 2, real(x), imag(x)))
program!(toks, x, k) = throw(Unsupported())
function program(f, nin)
    toks = SbTok[]
    if f isa CaptureArgs
        program!(toks, f, Ref(0))
    elseif haskey(FN, f)                              # map!(sin, ...), mapreduce(abs2, +, ...)
        foreach(i -> push!(toks, SbTok(0, i - 1, 0, 0)), 1:nin); push!(toks, SbTok(2, FN[f], 0, 0))
    else
        throw(Unsupported())                          # opaque closure: CPU fallback
    end
    toks
end
opcode(::Nothing) = 0
opcode(::Union{typeof(+),typeof(Base.add_sum)}) = 1
opcode(::Union{typeof(*),typeof(Base.mul_prod)}) = 2
opcode(::typeof(min)) = 3; opcode(::typeof(max)) = 4
opcode(op) = throw(Unsupported())
initcode(::Nothing) = (0, 0.0im); initcode(::typeof(zero)) = (1, 0.0im)
initcode(::typeof(identity)) = (2, 0.0im); initcode(::typeof(conj)) = (5, 0.0im)
initcode(f) = throw(Unsupported())   # `x -> β*x` / `x -> β` closures: pass them as ScaleBy(β) / ConstTo(β) from linalg.jl
struct ScaleBy{T}; β::T; end; initcode(s::ScaleBy) = (3, complex(s.β))
struct ConstTo{T}; β::T; end; initcode(s::ConstTo) = (4, complex(s.β))

pad(t::Tuple, n, x) = ntuple(i -> i <= length(t) ? t[i] : x, n)

# ---- the override: same signature as reference src/mapreduce.jl:98-99 ----------------------------------------
function _mapreduce_fuse!(f, op, initop, dims::Dims{N}, arrays::Tuple{Vararg{DeviceView}}) where {N}
    M = length(arrays)
    try
        (N <= SB_MAX_DIMS && M <= SB_MAX_OPS && all(a -> dtypecode(eltype(a)) >= 0, arrays)) || throw(Unsupported())
        toks = program(f, M - 1)
        length(toks) <= SB_MAX_TOKENS || throw(Unsupported())
        ic, β = initcode(initop)
        desc = SbDesc(N, M, pad(dims, SB_MAX_DIMS, 0),
            pad(map(a -> pad(strides(a), SB_MAX_DIMS, 0), arrays), SB_MAX_OPS, pad((), SB_MAX_DIMS, 0)),
            pad(map(a -> Ptr{Cvoid}(pointer(a.parent, a.offset + 1)), arrays), SB_MAX_OPS, C_NULL),   # mapreduce.jl:268
            pad(map(a -> Int32(dtypecode(eltype(a))), arrays), SB_MAX_OPS, Int32(0)),
            pad(map(a -> Int32(a.op === conj || a.op === adjoint), arrays), SB_MAX_OPS, Int32(0)),     # mapreduce.jl:276-278
            length(toks), pad(Tuple(toks), SB_MAX_TOKENS, SbTok(0, 0, 0, 0)), opcode(op), ic, real(β), imag(β))
        rc = (ALLREDUCE[] && op !== nothing) ? ccall((:sb_mapreduce_allreduce, LIB), Cint, (Ptr{Cvoid}, Ref{SbDesc}), ctx(), desc) :
             ccall((:sb_mapreduce, LIB), Cint, (Ptr{Cvoid}, Ref{SbDesc}), ctx(), desc)
        rc == 0 && return arrays[1]
        rc == -2 && throw(DimensionMismatch(unsafe_string(ccall((:sb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx()))))
        rc == -3 && throw(Unsupported())
        error(unsafe_string(ccall((:sb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
    catch e
        e isa Unsupported || rethrow()
        # CPU fallback: the untouched reference path on host copies (keeps "drop in unchanged" true for any f)
        host = map(a -> StridedView(tohost(a.parent), size(a), strides(a), a.offset, a.op), arrays)
        invoke(_mapreduce_fuse!, Tuple{Any,Any,Any,Dims,Tuple{Vararg{StridedView}}}, f, op, initop, dims, host)
        ccall((:sb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(),
              arrays[1].parent.ptr, host[1].parent, sizeof(host[1].parent))
        return arrays[1]
    end
end

# ---- reductions across GPUs (one Julia process per GPU) -------------------------------------------------------
# GPU analog of `threadedout` + serial fold (reference src/mapreduce.jl:153-170).  `allgather` is any function that
# returns the vector of every rank's 64-byte handle in rank order (MPI.Allgather, Distributed.jl, ...).
function attach_peers!(rank::Integer, world::Integer, allgather)
    h = Vector{UInt8}(undef, 64)
    ccall((:sb_peer_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx(), h) == 0 || error("sb_peer_export")
    all = reduce(vcat, allgather(h))
    ccall((:sb_peer_attach, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx(), rank, world, all) == 0 || error("sb_peer_attach")
end
const ALLREDUCE = Ref(false)   # `with_allreduce() do ... end`: reductions inside combine the partials of all ranks
with_allreduce(f) = (ALLREDUCE[] = true; try f() finally ALLREDUCE[] = false end)
# (inside `_mapreduce_fuse!` above, `sb_mapreduce` becomes `sb_mapreduce_allreduce` when ALLREDUCE[] && op !== nothing)

end # module
