"""Run one grouped batch (sb_mapreduce_batch: nprob same-plan permutedims of m^4 in ONE launch) a few times on cuda:0 --
for ncu / compute-sanitizer captures under gpurun.

    python tools/profile_batch.py [m=32] [nprob=8] [reps=6]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import strided_jl_b200 as sb  # noqa: E402


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    nprob = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    dev = torch.device("cuda", 0)
    sh = (m,) * 4
    col = (1, m, m * m, m ** 3)
    As = [torch.randn(m ** 4, dtype=torch.float64, device=dev) for _ in range(nprob)]
    Bs = [torch.zeros(m ** 4, dtype=torch.float64, device=dev) for _ in range(nprob)]
    calls = [([], 0, 0, 0.0, sh, [sb.StridedView(b, sh, col), sb.StridedView(a, sh, col).permutedims((3, 2, 1, 0))]) for a, b in zip(As, Bs)]
    for _ in range(reps):
        sb.run_batch(calls)
    torch.cuda.synchronize()
    for a, b in zip(As, Bs):
        assert torch.equal(b, a.view(*sh).permute(3, 2, 1, 0).contiguous().view(-1))
    print("batch done", sb.get_engine(0).stats())


if __name__ == "__main__":
    main()
