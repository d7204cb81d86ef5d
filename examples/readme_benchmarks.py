"""The four benchmark expressions of the reference's README (README.md:56-154), written against the host mirror of the
Strided.jl interface and run on cuda:0.  Julia on the left, this package on the right:

    @strided B .= (A .+ A') ./ 2                      B.assign((A + A.T) / 2)
    @strided B .= 3 .* A'                             B.assign(3 * A.T)
    @strided permutedims!(B, A, (4,3,2,1))            sb.permutedims_(B, A, (3, 2, 1, 0))
    @strided B .= permutedims(A,(1,2,3,4)) .+ ...     B.assign(((A + A.permutedims((1,2,3,0))) + ...) + ...)
    @strided mapreduce(abs2, +, A; dims=(2,3))        sb.mapreduce("abs2", "+", A, dims=(1, 2))

    python examples/readme_benchmarks.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import strided_jl_b200 as sb


def col(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def timed(fn, reps=20):
    """device time per call: a CUDA graph of `reps` calls, replayed"""
    eng = sb.get_engine(0)
    eng.set_sync(False)
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    eng.set_sync(True)
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    dev = "cuda"
    n = 4000
    a = torch.randn(n * n, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    A, B = sb.StridedView(a, (n, n), (1, n)), sb.StridedView(b, (n, n), (1, n))
    us = timed(lambda: B.assign((A + A.T) / 2))
    assert torch.equal(b.view(n, n), (a.view(n, n) + a.view(n, n).t()) * 0.5)
    print(f"B .= (A .+ A') ./ 2          4000x4000 Float64   {us:8.1f} us   (README @strided 4 threads: 30355 us)")

    n = 1000
    a = torch.randn(n * n, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    A, B = sb.StridedView(a, (n, n), (1, n)), sb.StridedView(b, (n, n), (1, n))
    us = timed(lambda: B.assign(3 * A.T))
    assert torch.equal(b.view(n, n), 3 * a.view(n, n).t())
    print(f"B .= 3 .* A'                 1000x1000 Float64   {us:8.1f} us   (README @strided 4 threads:   809 us)")

    m = 32
    shape = (m,) * 4
    a = torch.randn(m ** 4, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    A, B = sb.StridedView(a, shape, col(shape)), sb.StridedView(b, shape, col(shape))
    us = timed(lambda: sb.permutedims_(B, A, (3, 2, 1, 0)))
    # column-major parents: the torch view has its dims reversed
    assert torch.equal(b.view(*shape), a.view(*shape).permute(3, 2, 1, 0).contiguous())
    print(f"permutedims!(B,A,(4,3,2,1))  32^4 Float64        {us:8.1f} us   (README @strided 4 threads:  1192 us)")

    us = timed(lambda: B.assign(((A + A.permutedims((1, 2, 3, 0))) + A.permutedims((2, 3, 0, 1))) + A.permutedims((3, 0, 1, 2))))
    print(f"4-way permutedims sum        32^4 Float64        {us:8.1f} us   (README @strided 4 threads:  2796 us)")

    g, k = 8, 4096
    a = torch.randn(g * k * k, dtype=torch.float64, device=dev)
    A = sb.StridedView(a, (g, k, k), col((g, k, k)))
    out = sb.mapreduce("abs2", "+", A, dims=(1, 2))
    ref = (a.view(k * k, g) ** 2).sum(dim=0)
    assert torch.allclose(torch.as_tensor(out.to_numpy().reshape(-1), device=dev), ref, rtol=1e-10)
    print("mapreduce(abs2,+,A;dims=(2,3)) 8x4096x4096 Float64   ok   (the reference runs this on ONE task, src/mapreduce.jl:172-207)")


if __name__ == "__main__":
    main()
