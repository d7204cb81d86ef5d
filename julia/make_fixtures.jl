# make_fixtures.jl -- writes REFERENCE-PRODUCED golden fixtures for the parity suite.
#
# NOT EXECUTED in this repository (no Julia in the build image, SURVEY.md section 8c).  Run it on any machine that has
# Julia >= 1.6 with Strided.jl v2.3.2 (the reference commit 24fcf11), StridedViews.jl and NPZ.jl:
#
#     julia --project -e 'using Pkg; Pkg.add(["Strided", "StridedViews", "NPZ"])'
#     julia --project julia/make_fixtures.jl tests/golden_julia
#
# and commit the resulting tests/golden_julia/*.npz.  tests/test_golden_julia.py then replays every fixture through the
# C restatement of the reference (oracle/, CPU suite) and through the CUDA path (GPU suite) and compares with the output
# that the REAL Strided.jl produced -- this is what turns "parity unpinned" into "parity pinned" (DESIGN.md section 0 c).
#
# How a fixture is captured: the arrays of every recipe live in `RecArray`s, a DenseArray wrapper around a plain
# `Array`.  A method of `Strided._mapreduce_fuse!` for views over RecArrays (the same cut line the B200 glue uses,
# src/mapreduce.jl:98) records (f, op, initop, dims, views, parents before), runs the UNTOUCHED reference method on
# views over the underlying Arrays, and records the output parent afterwards.  The recipes themselves are written
# against the public API (`@strided`, `map!`, `permutedims!`, `mapreduce`, `mul!`), so the capture order of broadcast
# arguments (src/broadcast.jl:41-46), `promoteshape` zero strides (:56-65) and the adapters of src/mapreduce.jl:2-96 are
# all the reference's own.
#
# File format = tests/helpers.py `Case.to_npz` (one npz per `_mapreduce_fuse!` call):
#   nparents, parent0..parentK (flat, column-major), dims, vparent / voffset / vstrides / vconj (operand 0 = output),
#   tokens (ntok x 4: kind, a, re, im), op, initop, init_re, init_im, expected (output parent after the call), exact, rtol
using Strided, StridedViews, LinearAlgebra, Random, NPZ
include(joinpath(@__DIR__, "StridedB200.jl"))     # only `program`, `opcode`, `initcode` are used: nothing is ccall'ed
using .StridedB200: program, opcode, initcode, Unsupported

struct RecArray{T,N} <: DenseArray{T,N}
    data::Array{T,N}
end
Base.size(a::RecArray) = size(a.data)
Base.strides(a::RecArray) = strides(a.data)
Base.elsize(::Type{<:RecArray{T}}) where {T} = sizeof(T)
Base.getindex(a::RecArray, i::Int) = a.data[i]
Base.setindex!(a::RecArray, v, i::Int) = (a.data[i] = v; a)
Base.IndexStyle(::Type{<:RecArray}) = IndexLinear()
Base.pointer(a::RecArray, i::Integer=1) = pointer(a.data, i)
Base.unsafe_convert(::Type{Ptr{T}}, a::RecArray{T}) where {T} = Base.unsafe_convert(Ptr{T}, a.data)
Base.similar(a::RecArray, ::Type{T}, dims::Dims) where {T} = RecArray(Array{T}(undef, dims))
Base.reshape(a::RecArray, dims::Dims) = RecArray(reshape(a.data, dims))
rec(x::Array) = RecArray(x)

const RecView = StridedView{<:Any,<:Any,<:RecArray}
const RECORDS = Any[]
const CURRENT = Ref("")

function Strided._mapreduce_fuse!(f, op, initop, dims::Dims, arrays::Tuple{Vararg{RecView}})
    plain = map(a -> StridedView(a.parent.data, size(a), strides(a), a.offset, a.op), arrays)
    parents = Any[]                                    # distinct parents, by identity of the underlying Array
    pidx = map(plain) do v
        i = findfirst(p -> p === v.parent, parents)
        i === nothing ? (push!(parents, v.parent); length(parents)) : i
    end
    before = map(p -> copy(vec(p)), parents)
    toks = try program(f, length(arrays) - 1) catch e; e isa Unsupported ? nothing : rethrow() end
    codes = try (opcode(op), initcode(initop)) catch e; e isa Unsupported ? nothing : rethrow() end
    invoke(Strided._mapreduce_fuse!, Tuple{Any,Any,Any,Dims,Tuple{Vararg{StridedView}}}, f, op, initop, dims, plain)
    if toks !== nothing && codes !== nothing            # (recipes outside the program encoding are not fixtures)
        push!(RECORDS, (name=CURRENT[] * "_" * string(length(RECORDS)), dims=dims, views=plain, pidx=pidx, before=before,
                        toks=toks, op=codes[1], initop=codes[2][1], init=codes[2][2], expected=copy(vec(parents[pidx[1]]))))
    end
    return arrays[1]
end

function save(dir, r, exact)
    d = Dict{String,Any}()
    d["nparents"] = length(r.before)
    for (i, p) in enumerate(r.before)
        d["parent$(i - 1)"] = p
    end
    d["dims"] = collect(Int64, r.dims)
    d["vparent"] = collect(Int64, r.pidx .- 1)
    d["voffset"] = Int64[v.offset for v in r.views]
    d["vconj"] = Int64[(v.op === conj || v.op === adjoint) ? 1 : 0 for v in r.views]
    d["vstrides"] = permutedims(reduce(hcat, [collect(Int64, strides(v)) for v in r.views]))   # nviews x ndim
    d["tokens"] = isempty(r.toks) ? zeros(Float64, 0, 4) :
                  permutedims(reduce(hcat, [Float64[t.kind, t.a, t.re, t.im] for t in r.toks]))
    d["op"] = Int64(r.op); d["initop"] = Int64(r.initop)
    d["init_re"] = real(r.init); d["init_im"] = imag(r.init)
    d["expected"] = r.expected
    d["exact"] = Int64(exact)
    d["rtol"] = exact ? -1.0 : (real(eltype(r.expected)) == Float32 ? 1e-4 : 1e-10)
    npzwrite(joinpath(dir, r.name * ".npz"), d)
end

# ---- recipes: the BASELINE configs at test sizes + the recipes of test/othertests.jl ---------------------------------
function recipes()
    Random.seed!(1234)                                   # test/runtests.jl:7
    for T in (Float32, Float64, ComplexF32, ComplexF64)
        CURRENT[] = "c1_$(T)"; A = rec(randn(T, 100, 100)); B = rec(zeros(T, 100, 100)); @strided B .= 3 .* A'
        CURRENT[] = "c2_$(T)"; A = rec(randn(T, 96, 96)); B = rec(zeros(T, 96, 96)); @strided B .= (A .+ A') ./ 2
        CURRENT[] = "c3_$(T)"; A = rec(randn(T, 8, 8, 8, 8)); B = rec(zeros(T, 8, 8, 8, 8)); @strided permutedims!(B, A, (4, 3, 2, 1))
        CURRENT[] = "c4_$(T)"; A = rec(randn(T, 16, 16, 16, 16)); B = rec(zeros(T, 16, 16, 16, 16))
        @strided B .= permutedims(A, (1, 2, 3, 4)) .+ permutedims(A, (2, 3, 4, 1)) .+ permutedims(A, (3, 4, 1, 2)) .+ permutedims(A, (4, 1, 2, 3))
        CURRENT[] = "c5_$(T)"; A = rec(randn(T, 8, 64, 64)); @strided mapreduce(abs2, +, A; dims=(2, 3))
        CURRENT[] = "sum_$(T)"; A = rec(randn(T, 20011)); @strided sum(A)
        CURRENT[] = "adjoint_$(T)"; A = rec(randn(T, 37, 41)); B = rec(zeros(T, 41, 37)); @strided adjoint!(B, A)
        CURRENT[] = "conj_$(T)"; A = rec(randn(T, 37, 41)); @strided conj!(A)
        CURRENT[] = "axpby_$(T)"; X = rec(randn(T, 12, 12, 12)); Y = rec(randn(T, 12, 12, 12))
        axpby!(T(1) / 3, permutedims(StridedView(X), (3, 1, 2)), T(2) / 5, StridedView(Y))
        CURRENT[] = "lambda_$(T)"; A = rec(rand(T, 10, 10, 10)); B = rec(rand(T, 10, 10, 10)); C = rec(rand(T, 10, 10, 10)); D = rec(zeros(T, 10, 10, 10))
        @strided D .= sin.(A) .+ B ./ exp.(-abs.(C))                                    # test/othertests.jl:36
        CURRENT[] = "bcast_$(T)"; A = rec(randn(T, 10, 1)); B = rec(randn(T, 1, 10)); C = rec(zeros(T, 10, 10)); @strided C .= A .* B .+ 1
        CURRENT[] = "mul_generic_$(T)"; A = rec(randn(T, 7, 9)); B = rec(randn(T, 9, 5)); C = rec(randn(T, 7, 5))
        Strided.__mul!(StridedView(C), StridedView(A), StridedView(B), T(2), T(1) / 2)   # generic matmul client, src/linalg.jl:130-162
        CURRENT[] = "views_$(T)"; A = rec(randn(T, 30, 30)); B = rec(zeros(T, 10, 10))
        @strided B .= view(A, 1:3:30, 30:-3:1) .* 2                                     # stepped / reversed ranges (othertests.jl:141-181)
    end
end

function main()
    dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden_julia")
    mkpath(dir)
    recipes()
    for r in RECORDS
        # copy / permute / conj / adjoint and single-rounding maps are compared bit for bit; reductions and transcendentals by rtol
        transcend = any(t -> t.kind == 2 && t.a in (7, 8, 9, 10, 11, 12), r.toks)
        save(dir, r, r.op == 0 && !transcend && !(eltype(r.expected) <: Complex && any(t -> t.kind == 2 && t.a in (34, 35), r.toks)))
    end
    println("wrote $(length(RECORDS)) fixtures to $dir (Strided ", pkgversion(Strided), ")")
end
main()
