# StridedB200.jl -- the reference-side binding of libstrided_b200.so.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image nor on the GPU box (SURVEY.md
# section 8c).  It mirrors tests/helpers.py + strided.jl_b200/engine.py + strided.jl_b200/broadcast.py (`trace`) one to
# one, so every call below is exercised -- through ctypes instead of ccall -- by the GPU test-suite.
#
# What it does: adds methods for StridedViews whose parent lives in B200 HBM (`B200Array <: DenseArray`) at the single
# funnel of the hot path, `Strided._mapreduce_fuse!` (reference src/mapreduce.jl:98), so that every `@strided` block,
# `map!`, `permutedims!`, broadcast `copyto!`, `mapreduce` and the generic matmul `__mul!` (src/linalg.jl:130-162) keep
# working UNCHANGED -- nothing in Strided.jl is edited -- and land in `sb_mapreduce`.  The element function `f` and the
# `initop` arrive as Julia callables; they are turned into the postfix program of include/strided_b200.h either
# structurally (CaptureArgs trees, src/broadcast.jl:67-83; whitelisted functions) or by TRACING: the callable is run once
# on symbolic numbers (`Sym`) that record every operation, which covers the closures the reference itself passes
# (`x -> x * β`, `(x, y) -> x * y * α`, src/linalg.jl:149-158) and user lambdas built from the whitelisted functions.
# Anything the device path declines (SB_E_UNSUPPORTED, value-dependent control flow in a lambda, exotic eltypes) falls
# back to the original CPU method on host copies of the touched parents.
module StridedB200

using Strided, StridedViews, LinearAlgebra
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
        # background NVRTC compiles must have finished before the process runs its exit handlers (include/strided_b200.h)
        atexit(() -> ccall((:sb_shutdown, LIB), Cint, ()))
    end
    return CTX[]
end
lasterror() = unsafe_string(ccall((:sb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx()))

# ---- device-resident dense parent: what makes StridedView "device agnostic" (README.md:51-54) ------------
# A DenseArray of any rank (StridedView's parent must be a DenseArray; `similar` must return one, src/broadcast.jl:19-21).
# `owner` keeps the allocation alive for arrays that share it (reshape).
mutable struct B200Array{T,N} <: DenseArray{T,N}
    ptr::Ptr{T}
    dims::Dims{N}
    owner::Any
    function B200Array{T,N}(::UndefInitializer, dims::Dims{N}) where {T,N}
        p = Ref{Ptr{Cvoid}}()
        rc = ccall((:sb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}), ctx(), max(prod(dims), 1) * sizeof(T), p)
        rc == 0 || throw(OutOfMemoryError())
        v = new{T,N}(Ptr{T}(p[]), dims, nothing)
        finalizer(x -> ccall((:sb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.ptr), v)
    end
    B200Array{T,N}(ptr::Ptr{T}, dims::Dims{N}, owner) where {T,N} = new{T,N}(ptr, dims, owner)   # shares `owner`'s memory
end
B200Array{T}(::UndefInitializer, dims::Dims{N}) where {T,N} = B200Array{T,N}(undef, dims)
B200Array{T}(::UndefInitializer, dims::Integer...) where {T} = B200Array{T}(undef, Dims(dims))
const B200Vector{T} = B200Array{T,1}
Base.size(a::B200Array) = a.dims
Base.strides(a::B200Array) = Base.size_to_strides(1, a.dims...)
Base.elsize(::Type{<:B200Array{T}}) where {T} = sizeof(T)
Base.pointer(a::B200Array{T}, i::Integer=1) where {T} = a.ptr + (i - 1) * sizeof(T)
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where {T} = a.ptr
Base.similar(a::B200Array, ::Type{T}, dims::Dims{N}) where {T,N} = B200Array{T,N}(undef, dims)
Base.reshape(a::B200Array{T}, dims::Dims{N}) where {T,N} =
    (prod(dims) == length(a) || throw(DimensionMismatch("reshape")); B200Array{T,N}(a.ptr, dims, something(a.owner, a)))
# scalar access is for debugging / show only (one 8-byte copy over PCIe per element)
function Base.getindex(a::B200Array{T}, i::Int) where {T}
    @boundscheck checkbounds(a, i)
    r = Ref{T}()
    ccall((:sb_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), r, pointer(a, i), sizeof(T))
    return r[]
end
function Base.setindex!(a::B200Array{T}, v, i::Int) where {T}
    @boundscheck checkbounds(a, i)
    r = Ref{T}(convert(T, v))
    ccall((:sb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), pointer(a, i), r, sizeof(T))
    return a
end
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
function todevice(a::Array{T,N}) where {T,N}
    v = B200Array{T,N}(undef, size(a))
    ccall((:sb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), v.ptr, a, sizeof(a)) == 0 || error(lasterror())
    return v
end
function tohost(v::B200Array{T,N}) where {T,N}
    a = Array{T,N}(undef, size(v))
    ccall((:sb_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), a, v.ptr, sizeof(a)) == 0 || error(lasterror())
    return a
end
Base.Array(v::B200Array) = tohost(v)

const DeviceView{T,N} = StridedView{T,N,<:B200Array}

# ---- (f, op, initop) -> program / codes -------------------------------------------------------------------
dtypecode(::Type{Float32}) = 0; dtypecode(::Type{Float64}) = 1
dtypecode(::Type{ComplexF32}) = 2; dtypecode(::Type{ComplexF64}) = 3
dtypecode(::Type) = -1
const FN1 = Dict{Any,Int32}(identity => 0, (-) => 1, conj => 2, abs => 3, abs2 => 4, real => 5, imag => 6, sqrt => 7,
    exp => 8, log => 9, sin => 10, cos => 11, tanh => 12, inv => 13)
const FN2 = Dict{Any,Int32}((+) => 32, (-) => 33, (*) => 34, (/) => 35, max => 36, min => 37, (<) => 38)
struct Unsupported <: Exception end

constok(x::Union{Integer,Rational,Bool}) = SbTok(1, 0, Float64(x), 0)        # weak literal (`3`, `1//2`): adopts the operand type
constok(x::Float32) = SbTok(1, 1, x, 0)
constok(x::Float64) = SbTok(1, 2, x, 0)
constok(x::Complex) = SbTok(1, real(x) isa Float32 ? 1 : 2, real(x), imag(x))
constok(x) = throw(Unsupported())

# ---- tracing: run a callable once on symbolic numbers that record the operations --------------------------------
# Same idea as strided.jl_b200/broadcast.py `trace`.  Value-dependent control flow (`x < 0 ? ... : ...`) needs a Bool and
# therefore throws -> Unsupported -> CPU fallback, which is the correct answer for such lambdas.
struct Sym <: Number
    toks::Vector{SbTok}
end
Sym(x::Number) = Sym([constok(x)])
Base.promote_rule(::Type{Sym}, ::Type{<:Number}) = Sym
Base.convert(::Type{Sym}, x::Sym) = x
Base.convert(::Type{Sym}, x::Number) = Sym(x)
for (f, id) in ((:-, 1), (:conj, 2), (:abs, 3), (:abs2, 4), (:real, 5), (:imag, 6), (:sqrt, 7), (:exp, 8), (:log, 9),
                (:sin, 10), (:cos, 11), (:tanh, 12), (:inv, 13))
    @eval Base.$f(x::Sym) = Sym(vcat(x.toks, SbTok(2, $id, 0, 0)))
end
Base.:+(x::Sym) = x
Base.identity(x::Sym) = x
Base.zero(::Sym) = Sym(0); Base.one(::Sym) = Sym(1)
Base.zero(::Type{Sym}) = Sym(0); Base.one(::Type{Sym}) = Sym(1)
Base.adjoint(x::Sym) = conj(x); Base.transpose(x::Sym) = x
for (f, id) in ((:+, 32), (:-, 33), (:*, 34), (:/, 35), (:max, 36), (:min, 37))
    @eval Base.$f(x::Sym, y::Sym) = Sym(vcat(x.toks, y.toks, SbTok(2, $id, 0, 0)))
end
Base.:<(x::Sym, y::Sym) = throw(Unsupported())     # a Bool result would steer control flow: not traceable
Base.:(==)(x::Sym, y::Sym) = throw(Unsupported())
Base.isless(x::Sym, y::Sym) = throw(Unsupported())
function tracecall(f, nin::Int)
    try
        r = f(ntuple(i -> Sym([SbTok(0, i - 1, 0, 0)]), nin)...)
        r isa Sym || (r = Sym(r))           # constant function
        return r.toks
    catch e
        (e isa Unsupported || e isa MethodError || e isa TypeError || e isa InexactError) && throw(Unsupported())
        rethrow()
    end
end

# depth-first, left-to-right: exactly the order in which `consume` (broadcast.jl:86-98) eats the arguments
function program!(toks, c::CaptureArgs, k::Ref{Int})
    nargs = length(c.args)
    f = c.f
    if (f === +) || (f === *)                       # n-ary: left fold, as Julia evaluates it
        program!(toks, c.args[1], k)
        for a in c.args[2:end]
            program!(toks, a, k); push!(toks, SbTok(2, FN2[f], 0, 0))
        end
        return
    end
    if (nargs == 1 && haskey(FN1, f)) || (nargs == 2 && haskey(FN2, f))
        foreach(a -> program!(toks, a, k), c.args)
        push!(toks, SbTok(2, nargs == 1 ? FN1[f] : FN2[f], 0, 0))
        return
    end
    # an arbitrary callable inside the broadcast tree (`f.(A, B)`): trace it over sub-programs of its arguments
    subs = map(c.args) do a
        t = SbTok[]; program!(t, a, k); Sym(t)
    end
    r = try f(subs...) catch e; (e isa Unsupported || e isa MethodError || e isa TypeError) ? throw(Unsupported()) : rethrow() end
    r isa Sym || (r = Sym(r))
    append!(toks, r.toks)
end
program!(toks, ::Arg, k::Ref{Int}) = (push!(toks, SbTok(0, k[], 0, 0)); k[] += 1)
program!(toks, x::Number, k) = push!(toks, constok(x))
program!(toks, x, k) = throw(Unsupported())
function program(f, nin)
    toks = SbTok[]
    if f isa CaptureArgs
        program!(toks, f, Ref(0))
    elseif nin == 1 && haskey(FN1, f)                 # map!(sin, ...), mapreduce(abs2, +, ...)
        push!(toks, SbTok(0, 0, 0, 0)); push!(toks, SbTok(2, FN1[f], 0, 0))
    elseif nin == 2 && haskey(FN2, f)                 # _mapreducedim!(*, +, ...) of __mul!  (linalg.jl:145-149)
        push!(toks, SbTok(0, 0, 0, 0)); push!(toks, SbTok(0, 1, 0, 0)); push!(toks, SbTok(2, FN2[f], 0, 0))
    elseif (f === +) || (f === *)
        push!(toks, SbTok(0, 0, 0, 0))
        for i in 2:nin
            push!(toks, SbTok(0, i - 1, 0, 0)); push!(toks, SbTok(2, FN2[f], 0, 0))
        end
    else
        toks = tracecall(f, nin)                      # closures: `(x, y) -> x * y * α` (linalg.jl:152), user lambdas
    end
    toks
end
opcode(::Nothing) = 0
opcode(::Union{typeof(+),typeof(Base.add_sum)}) = 1
opcode(::Union{typeof(*),typeof(Base.mul_prod)}) = 2
opcode(::typeof(min)) = 3; opcode(::typeof(max)) = 4
opcode(op) = throw(Unsupported())     # `&` / `|` (mapreduce.jl:186-187) reduce Bool arrays: not a device eltype -> CPU method

# initop -> (code, β): traced like f and matched against the flavours of the ABI (SB_INIT_*)
initcode(::Nothing) = (0, 0.0im); initcode(::typeof(zero)) = (1, 0.0im)
initcode(::typeof(identity)) = (2, 0.0im); initcode(::typeof(conj)) = (5, 0.0im)
function initcode(f)                  # `x -> x * β` (linalg.jl:149,158), `x -> β * x`, `x -> β`, `x -> zero(x)` ...
    t = tracecall(f, 1)
    isarg(k) = k.kind == 0
    iscst(k) = k.kind == 1
    length(t) == 1 && isarg(t[1]) && return (2, 0.0im)
    length(t) == 1 && iscst(t[1]) && return (4, complex(t[1].re, t[1].im))
    length(t) == 2 && isarg(t[1]) && t[2].kind == 2 && t[2].a == 2 && return (5, 0.0im)
    if length(t) == 3 && t[3].kind == 2 && t[3].a == 34
        isarg(t[1]) && iscst(t[2]) && return (3, complex(t[2].re, t[2].im))
        iscst(t[1]) && isarg(t[2]) && return (3, complex(t[1].re, t[1].im))
    end
    throw(Unsupported())
end

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
        rc == -2 && throw(DimensionMismatch(lasterror()))
        rc == -3 && throw(Unsupported())
        error(lasterror())
    catch e
        e isa Unsupported || rethrow()
        return cpu_fallback!(f, op, initop, dims, arrays)
    end
end

# CPU fallback: the untouched reference path on host copies (keeps "drop in unchanged" true for any f).  Every DISTINCT
# parent is downloaded once (aliased views share the copy), only the output's parent travels back.
function cpu_fallback!(f, op, initop, dims, arrays)
    hosts = IdDict{Any,Any}()
    hostview(a) = StridedView(get!(() -> tohost(a.parent), hosts, a.parent), size(a), strides(a), a.offset, a.op)
    host = map(hostview, arrays)
    invoke(_mapreduce_fuse!, Tuple{Any,Any,Any,Dims,Tuple{Vararg{StridedView}}}, f, op, initop, dims, host)
    out = hosts[arrays[1].parent]
    ccall((:sb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), arrays[1].parent.ptr, out, sizeof(out)) == 0 || error(lasterror())
    return arrays[1]
end

# ---- reductions across GPUs (one Julia process per GPU) -------------------------------------------------------
# GPU analog of `threadedout` + serial fold (reference src/mapreduce.jl:153-170).  `allgather` is any function that
# returns the vector of every rank's 64-byte handle in rank order (MPI.Allgather, Distributed.jl, ...); `barrier` any
# function that returns once all ranks have called it (no rank may push into a buffer that is still being attached).
function attach_peers!(rank::Integer, world::Integer, allgather, barrier=() -> nothing)
    h = Vector{UInt8}(undef, 64)
    ccall((:sb_peer_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx(), h) == 0 || error("sb_peer_export")
    all = reduce(vcat, allgather(h))
    ccall((:sb_peer_attach, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx(), rank, world, all) == 0 || error("sb_peer_attach")
    barrier()
end
const ALLREDUCE = Ref(false)   # `with_allreduce() do ... end`: reductions inside combine the partials of all ranks
with_allreduce(f) = (ALLREDUCE[] = true; try f() finally ALLREDUCE[] = false end)
# (inside `_mapreduce_fuse!` above, `sb_mapreduce` becomes `sb_mapreduce_allreduce` when ALLREDUCE[] && op !== nothing)

end # module
