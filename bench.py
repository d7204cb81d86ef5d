#!/usr/bin/env python
"""bench.py -- headline benchmark of the strided map/permute/reduce hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--no-configs] [--no-sharded]

Headline workload (BASELINE.json configs[1]): Float64 4000x4000  `@strided B .= (A .+ A') ./ 2`.
A "step" is one pass of the hot path over one such problem per GPU (N GPUs = N independent problems, i.e. an
outer batch dimension sharded over the ranks: weak scaling, no data-path collective).
Metric: effective GB/s = ALGORITHMIC bytes / time, algorithmic bytes = compulsory traffic = 8 B x (distinct
input elements + output elements) = 256 000 000 B per problem (A and A' alias the same memory).

  value     device-resident throughput, K steps timed with CUDA events on the launching stream, max over ranks
  e2e       same metric through the C ABI with HOST buffers (sb_mapreduce_host), all copies inside the timed region
  roofline  the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the restated reference CPU path (oracle/, all host threads) on the same workload
  configs   (N = 1) every other BASELINE config, device-timed in the same run: C1, C3, C4, C4', C5 and the per-GPU
            share of C5 -- ms, GB/s, fraction of the HBM peak, kernel family
  sharded   BASELINE configs[4] as specified: ONE Float64 8x4096x4096 `mapreduce(abs2,+,A;dims=(2,3))` split on dim 1
            over the N ranks (dense slab per GPU, no collective), timed max-over-ranks -> strong scaling; plus the
            complete reduction sum(abs2, A) with the ranks' partials exchanged through NVLink peer memory INSIDE the
            reduction kernel (sb_mapreduce_allreduce) next to local kernel + NCCL all-reduce.  Every result is
            asserted against torch / math.fsum in the run.

`--impl reference` times the reference's own CPU path: Julia is not in this image, so this is the C
restatement of Strided.jl's planner + task bisection + blocked loop nest (oracle/strided_ref.c), kind "port".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_MAT = 4000
ALG_BYTES = 2 * N_MAT * N_MAT * 8  # read A once + write B once
METRIC = "effective_GBps_strided_map_f64_4000x4000_A_plus_At_over_2"
WORKLOAD = "BASELINE configs[1]: Float64 4000x4000 B .= (A .+ A')./2"  # identical in both arms
PUBLISHED_GBPS = 8.43  # README.md:120-121, `@strided` 4 threads, 30.355 ms -> 256 MB / 30.355 ms (hardware not stated)
TOKENS = [(0, 0, 0.0, 0.0), (0, 1, 0.0, 0.0), (2, 32, 0.0, 0.0), (1, 0, 2.0, 0.0), (2, 35, 0.0, 0.0)]  # (A + A') / 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()

    def summary(self):
        try:
            self.f.flush()
            rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
            os.unlink(self.f.name)
            sm = sorted(float(r[1]) for r in rows)
            mx = max(float(r[2]) for r in rows)
            reasons = set()
            for r in rows:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(rows)}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_reference(budget_s, min_reps=3, nthreads=None):
    """The restated reference CPU path (oracle/strided_ref.c) on the full C2 problem, all host threads (or `nthreads`)."""
    import strided_jl_b200 as sb
    from oracle import ref as oref
    nthreads = nthreads or host_threads()
    rng = np.random.default_rng(1234)
    a = rng.standard_normal(N_MAT * N_MAT)
    b = np.zeros_like(a)
    A = sb.StridedView(a, (N_MAT, N_MAT), (1, N_MAT))
    B = sb.StridedView(b, (N_MAT, N_MAT), (1, N_MAT))
    desc = sb.make_desc(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), [B, A, A.T])
    oref.mapreduce(desc, nthreads)  # warm-up (page faults, thread pool)
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_reps or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        oref.mapreduce(desc, nthreads)
        times.append(time.perf_counter() - t0)
        if len(times) >= 2000:
            break
    # sanity: the timed thing computed the right answer
    am = a.reshape(N_MAT, N_MAT)
    assert np.array_equal(b.reshape(N_MAT, N_MAT)[:64, :64], ((am + am.T) / 2)[:64, :64])
    return times, nthreads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = one bounded sample: ~budget seconds of repeated full-problem passes
    per_step_budget = max(0.2, min(2.0, 20.0 / max(1, args.steps + args.warmup)))
    all_times = []
    nthreads = host_threads()
    for s in range(args.warmup + args.steps):
        times, nthreads = time_reference(per_step_budget, min_reps=2)
        if s >= args.warmup:
            all_times.extend(times)
    ms = float(np.mean(all_times)) * 1e3
    val = ALG_BYTES / (ms * 1e-3) / 1e9
    t1, _ = time_reference(1.0, min_reps=2, nthreads=1)  # the same path on ONE task (README.md:72-73 publishes 56.2 ms)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": val / PUBLISHED_GBPS,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algorithmic_bytes": ALG_BYTES},
        "note": "restated Strided.jl CPU path (oracle/strided_ref.c): Julia is not installed in this image",
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": nthreads, "kind": "port",
                         "sample": f"{len(all_times)} full passes of the 4000x4000 problem, mean; min {min(all_times) * 1e3:.2f} ms",
                         "value_1_thread": ALG_BYTES / float(np.mean(t1)) / 1e9},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def _graph_time(torch, fn, reps, stream, barrier=None, replays=3):
    """ms per call: `reps` calls captured in one CUDA graph on `stream`, best of `replays` replays (device time)."""
    with torch.cuda.stream(stream):
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(reps):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(replays):
            if barrier is not None:
                barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
    del g
    return best


def sharded_record(torch, sb, dist, eng, rank, world, dev, peak):
    """BASELINE configs[4]: ONE Float64 8x4096x4096 mapreduce(abs2,+,A;dims=(2,3)) split on dim 1 over the ranks.
    Rank r holds slices [r*per, (r+1)*per) as DENSE 4096x4096 slabs in its own HBM and produces out[slices]: no
    collective on the data path (reference: mapreduce.jl:16-30, :74-96; the reference itself runs this config on one
    task, :203-207).  Strong scaling: the problem size is fixed, `t1` (all 8 slices on one GPU, same placement) is
    measured by rank 0 in the same run.  Then sum(abs2, A), the complete reduction (:153-170 analog): per-rank partial,
    exchange of ONE element per rank, fold in rank order."""
    from strided_jl_b200 import sharded
    G, K = 8, 4096
    if G % world != 0:
        return {"skipped": f"8 slices do not split evenly over {world} ranks"}
    per = G // world
    prog = [(0, 0, 0.0, 0.0), (2, sb.abi.FN["abs2"], 0.0, 0.0)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    slab = torch.randn(per * K * K, dtype=torch.float64, device=dev, generator=gen)
    A = sb.StridedView(slab, (per, K, K), (K * K, 1, K))
    out = torch.zeros(per, dtype=torch.float64, device=dev)
    O = sb.StridedView(out, (per, K, K), (1, 0, 0))
    eng.set_sync(False)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        sb.run_mapreduce(prog, 1, 1, 0.0, (per, K, K), [O, A])  # initop = zero: out[g] = sum_jk A[g,j,k]^2

    side = torch.cuda.Stream(device=dev)
    ms = _graph_time(torch, step, 40, side, barrier)
    want = (slab.view(per, K * K) ** 2).sum(dim=1)
    assert torch.allclose(out, want, rtol=1e-12, atol=0.0), ("sharded C5: per-slice sums differ from torch", out, want)
    from bench_configs import _kernel
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_bytes = G * K * K * 8 + G * 8
    rec = {
        "workload": "BASELINE configs[4]: Float64 8x4096x4096 mapreduce(abs2,+,A;dims=(2,3)), ONE problem sharded on dim 1",
        "n_gpus": world, "slices_per_gpu": per, "scaling": "strong", "algorithmic_bytes": total_bytes,
        "us_per_step_max_over_ranks": ms * 1e3, "aggregate_GBps": total_bytes / (ms * 1e-3) / 1e9,
        "frac_of_N_x_peak": total_bytes / (ms * 1e-3) / 1e9 / (peak * world),
        "kernel": _kernel(prog, 1, 1, (per, K, K), [O, A]),
        "placement": "dense 4096x4096 slab per slice in each GPU's HBM; no data-path collective",
        "checked": "every out[g] against torch (rtol 1e-12) on every rank",
    }
    # t1: the whole problem on ONE GPU with the same placement (rank 0 only; the other ranks wait at the barrier)
    if world > 1:
        t1 = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank == 0:
            full = torch.randn(G * K * K, dtype=torch.float64, device=dev, generator=gen)
            A1 = sb.StridedView(full, (G, K, K), (K * K, 1, K))
            o1 = torch.zeros(G, dtype=torch.float64, device=dev)
            O1 = sb.StridedView(o1, (G, K, K), (1, 0, 0))
            t1[0] = _graph_time(torch, lambda i: sb.run_mapreduce(prog, 1, 1, 0.0, (G, K, K), [O1, A1]), 20, side)
            assert torch.allclose(o1, (full.view(G, K * K) ** 2).sum(dim=1), rtol=1e-12, atol=0.0)
            del full
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        rec["us_one_gpu_same_run"] = float(t1.item()) * 1e3
        rec["strong_scaling_speedup"] = float(t1.item()) / ms
        rec["strong_scaling_efficiency"] = float(t1.item()) / ms / world
    # ---- complete reduction: one element per rank is exchanged -------------------------------------------------
    if world > 1:
        sharded.attach_peer_group()
        total = torch.zeros(1, dtype=torch.float64, device=dev)
        T = sb.StridedView(total, (per, K, K), (0, 0, 0))

        def fstep(i):  # ONE launch per rank: local reduction + exchange through peer memory + fold in rank order
            sb.run_mapreduce(prog, 1, 1, 0.0, (per, K, K), [T, A], allreduce=True)

        launches0 = eng.stats()["launches"]
        fstep(0)
        torch.cuda.synchronize()
        launches_per_call = eng.stats()["launches"] - launches0
        # oracle for the total: math.fsum over the ranks' exact per-slice sums gathered on the host
        parts = [None] * world
        dist.all_gather_object(parts, [float(x) for x in (slab.view(per, K * K) ** 2).sum(dim=1).cpu()])
        want_total = math.fsum(x for p in parts for x in p)
        got = float(total.item())
        assert abs(got - want_total) <= 1e-11 * abs(want_total), ("fused peer exchange: total differs from math.fsum", got, want_total)
        same = [None] * world
        dist.all_gather_object(same, got)
        assert all(s == same[0] for s in same), ("ranks disagree on the folded total (must be bit-identical)", same)
        ms_fused = _graph_time(torch, fstep, 40, side, barrier)
        # the graph replays ran on fresh epochs: the total is still right after ~160 replayed collective calls
        assert float(total.item()) == got
        tf = torch.tensor([ms_fused], dtype=torch.float64, device=dev)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        # reference point: the same step with the exchange done by NCCL (local reduction kernel + all_reduce of one
        # element); launched eagerly (host-paced) and timed with events -- NCCL is not captured into a graph here
        tot2 = torch.zeros(1, dtype=torch.float64, device=dev)
        T2 = sb.StridedView(tot2, (per, K, K), (0, 0, 0))

        def nstep():
            sb.run_mapreduce(prog, 1, 1, 0.0, (per, K, K), [T2, A])
            dist.all_reduce(tot2)

        for _ in range(3):
            nstep()
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(20):
            nstep()
        n1.record()
        torch.cuda.synchronize()
        assert abs(float(tot2.item()) - want_total) <= 1e-11 * abs(want_total)
        tn = torch.tensor([n0.elapsed_time(n1) / 20], dtype=torch.float64, device=dev)
        # the fused call launched eagerly the same way (like for like with the NCCL line)
        barrier()
        n0.record()
        for i in range(20):
            fstep(i)
        n1.record()
        torch.cuda.synchronize()
        te = torch.tensor([n0.elapsed_time(n1) / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        eng.peer_detach()
        rec["complete_reduction"] = {
            "what": "sum(abs2, A) over the whole 1 GiB array: per-rank partial + exchange of ONE element per rank + fold in rank order",
            "fused_peer_exchange_us": float(tf.item()) * 1e3, "fused_launches_per_call": launches_per_call,
            "fused_aggregate_GBps": total_bytes / (float(tf.item()) * 1e-3) / 1e9,
            "fused_eager_us": float(te.item()) * 1e3, "local_kernel_plus_nccl_allreduce_eager_us": float(tn.item()) * 1e3,
            "exchange_cost_over_local_us": (float(tf.item()) - ms) * 1e3,
            "timing": "fused: 40 collective calls captured in one CUDA graph, replayed (epochs advance on the device); eager lines are host-paced",
            "checked": "total against math.fsum of the per-slice sums (1e-11 relative); bit-identical on all ranks; NCCL variant against the same oracle",
        }
    eng.set_sync(True)
    del slab
    return rec


def copy_ceiling(torch, dist, dev, a_host, b_host, barrier):
    """What the host link of this box gives when NOTHING but copies runs: 128 MB H2D and 128 MB D2H on two streams at the
    same time, on every rank at once (max over ranks) -- the ceiling for e2e, to separate the box from the library."""
    d_in = torch.empty(N_MAT * N_MAT, dtype=torch.float64, device=dev)
    d_out = torch.zeros(N_MAT * N_MAT, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    reps = 4
    for timed in (False, True):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(a_host, non_blocking=True)
            with torch.cuda.stream(s2):
                b_host.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


def run_ours(args):
    import torch
    import strided_jl_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = sb.get_engine(local_rank)
    peak, peak_src = peaks()

    rng = np.random.default_rng(1234 + rank)
    a_host = torch.from_numpy(rng.standard_normal(N_MAT * N_MAT)).pin_memory()
    b_host = torch.empty(N_MAT * N_MAT, dtype=torch.float64).pin_memory()
    a = a_host.to(dev, non_blocking=False)
    b = torch.empty_like(a)
    A = sb.StridedView(a, (N_MAT, N_MAT), (1, N_MAT))
    B = sb.StridedView(b, (N_MAT, N_MAT), (1, N_MAT))
    expr_views = [B, A, A.T]

    def step():
        sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), expr_views)

    eng.set_sync(False)  # stream-ordered launches; the timed region is bracketed by synchronize()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    # correctness of what is being timed (sampled)
    am = a.view(N_MAT, N_MAT)
    assert torch.equal(b.view(N_MAT, N_MAT)[:128, :128], ((am + am.t()) * 0.5)[:128, :128])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # The K timed steps are captured once into a CUDA graph and replayed: the number is device time of K back-to-back
    # launches, free of the Python/ctypes cost of building K descriptors (~15 us per call, comparable to the 44 us
    # kernel).  The same K steps launched eagerly from Python are timed too and reported as `eager_ms_per_step`.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = torch.cuda.Stream(device=dev)
    with ClockSampler(local_rank) as cs:
        with torch.cuda.stream(side):
            step()  # binds the engine to this stream (plan cached)
            torch.cuda.synchronize()
            eng.reset_stats()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(args.steps):
                    step()
            launches_timed = eng.stats()["launches"]  # kernels of libstrided_b200.so inside the timed region (= graph nodes)
            graph.replay()
            barrier()
            ev0.record(side)
            graph.replay()
            ev1.record(side)
            barrier()
            elapsed_ms = ev0.elapsed_time(ev1)
        # eager launches of the same K steps (host-paced)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        eager_ms = e0.elapsed_time(e1) / args.steps
        # keep the sampler alive for a minimum window so that short runs still get clock samples under load
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            graph.replay()
        torch.cuda.synchronize()
    clocks = cs.summary()
    clocks["window"] = "timed graph replay + eager steps + 0.5 s of replays of the same graph (the timed region alone is shorter than one nvidia-smi sample)"
    del graph
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * ALG_BYTES / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    # Every step: the input A comes from pinned HOST memory and the result B lands in pinned HOST memory, all through
    # sb_mapreduce_host, all inside the timed region.  Two contexts (two streams) are used alternately in stream-ordered
    # mode, so that consecutive steps overlap on the full-duplex host link; every step still moves all of its bytes.
    eng.set_sync(True)
    engs = [sb.engine.Engine(local_rank), sb.engine.Engine(local_rank)]
    b_hosts = [b_host, torch.empty(N_MAT * N_MAT, dtype=torch.float64).pin_memory()]
    Ah = sb.StridedView(a_host.numpy(), (N_MAT, N_MAT), (1, N_MAT))
    host_views = [[sb.StridedView(bh.numpy(), (N_MAT, N_MAT), (1, N_MAT)), Ah, Ah.T] for bh in b_hosts]
    e2e_steps = max(4, min(args.steps, 20))
    for e in engs:
        e.set_sync(False)
    for i in range(4):
        sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), host_views[i % 2], engine=engs[i % 2])
    for e in engs:
        e.synchronize()
        e.reset_stats()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), host_views[i % 2], engine=engs[i % 2])
    for e in engs:
        e.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    est = {k: sum(e.stats()[k] for e in engs) for k in ("h2d_bytes", "d2h_bytes", "launches")}
    for bh in b_hosts:
        assert np.array_equal(bh.numpy()[:4096], b.cpu().numpy()[:4096])
        assert np.array_equal(bh.numpy()[-4096:], b.cpu().numpy()[-4096:])
    e2e_value = world * ALG_BYTES / (e2e_s / e2e_steps) / 1e9
    # the same call made synchronously, one at a time (latency view)
    engs[0].set_sync(True)
    b_hosts[0].zero_()
    sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), host_views[0], engine=engs[0])  # (plan of the synchronous mode, untimed)
    assert np.array_equal(b_hosts[0].numpy()[:4096], b.cpu().numpy()[:4096]) and np.array_equal(b_hosts[0].numpy()[-4096:], b.cpu().numpy()[-4096:])
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), host_views[0], engine=engs[0])
    e2e_sync_s = (time.perf_counter() - t0) / 5
    if dist is not None:
        t = torch.tensor([e2e_sync_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sync_s = float(t.item())
    e2e_sync_value = world * ALG_BYTES / e2e_sync_s / 1e9
    zc_calls = engs[0].stats().get("zero_copy_calls", 0)
    ceil_s = copy_ceiling(torch, dist, dev, a_host, b_host, barrier)
    for e in engs:
        e.close()

    # ---- BASELINE configs[4] sharded over the ranks + fused peer exchange (asserted) ---------------------------------
    sharded = None
    if not args.no_sharded:
        try:
            sharded = sharded_record(torch, sb, dist, eng, rank, world, dev, peak)
        except AssertionError:
            raise
        except Exception as e:  # e.g. CUDA IPC unavailable on this box: report, do not lose the headline
            sharded = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        achieved = ALG_BYTES / (ms_per_step * 1e-3) / 1e9  # per GPU: one launch per step
        traffic = None
        prof = os.path.join(ROOT, "profiles", "c2_kernel_ncu.json")  # written by tools/summarize_ncu.py from an `ncu --set full` capture
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        plan = sb.plan_describe(sb.make_desc(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), expr_views))
        if plan.get("orbit"):  # alias-fused path: A read once by TMA, both views served from shared memory, TMA store
            kernel = ("map_orbit_kernel<double, add2_mul, NIN=2, EPT=%d> (alias-fused orbits, %d-stage TMA ring, TMA store)"
                      % (plan["orbit"]["ept"], plan["orbit"]["nstage"]))
        elif plan.get("tma"):
            kernel = "map_tma_kernel<double, add2_mul, NIN=2, EPT=8> (TMA ring, %d stages)" % plan["tma"]
        else:
            kernel = "map_tile_kernel<double, add2_mul, NIN=2, EPT=8>"
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "eager_ms_per_step": eager_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": value / PUBLISHED_GBPS,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "algorithmic_bytes": ALG_BYTES},
            "details": {"algorithmic_bytes_per_gpu": ALG_BYTES, "operand_bytes_per_gpu": 3 * N_MAT * N_MAT * 8,
                        "l2": "working set 256 MB (A 128 MB + B 128 MB) exceeds the 126 MB L2; no explicit flush",
                        "parallelism": f"{world} independent problems, one per GPU (batch dim sharded), no collective",
                        "timing": "K steps captured in one CUDA graph, one replay timed with CUDA events on the launching stream",
                        "vs_baseline_ref": "README.md:120-121 @strided 4 threads 30.355 ms = 8.43 GB/s, hardware not stated",
                        "plan": plan},
            "gpu_launches": launches_timed,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": est["h2d_bytes"] // e2e_steps,
                    "d2h_bytes_per_step": est["d2h_bytes"] // e2e_steps, "steps": e2e_steps,
                    "api": "sb_mapreduce_host (C ABI, pinned host buffers), 2 contexts alternating, stream-ordered: staged "
                           "(H2D copy of A, kernel, D2H copy of B per step; consecutive steps overlap on the full-duplex link)",
                    "one_call_at_a_time": e2e_sync_value,
                    "one_call_at_a_time_mode": ("zero-copy: ONE kernel reads A from pinned host memory (once: alias-fused orbit plan) and writes B "
                                                "to pinned host memory, both link directions overlap inside the call" if zc_calls else
                                                "staged: H2D copy, kernel, D2H copy, one after the other"),
                    "copy_only_ceiling": world * ALG_BYTES / ceil_s / 1e9,
                    "copy_only_ceiling_note": "cudaMemcpyAsync of 128 MB H2D and 128 MB D2H concurrently on two streams, all ranks at once, "
                                              "no kernel: what this box's host link gives, in the same unit"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "static: profiles/c2_kernel_ncu.json (one `ncu --set full` capture, per launch), not measured in this run",
                         "kernel": kernel, "peak_source": peak_src,
                         "basis": "algorithmic bytes 256e6 per launch / CUDA-event time per launch"},
        }
        if sharded is not None:
            line["sharded"] = sharded
        if world == 1:
            times, nthreads = time_reference(budget_s=10.0)
            cv = ALG_BYTES / float(np.mean(times)) / 1e9
            t1, _ = time_reference(1.5, min_reps=2, nthreads=1)
            line["cpu_baseline"] = {"value": cv, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                    "value_1_thread": ALG_BYTES / float(np.mean(t1)) / 1e9,
                                    "sample": f"{len(times)} full passes of the same 4000x4000 problem (restated Strided.jl CPU path, "
                                              f"{nthreads} tasks), mean {np.mean(times) * 1e3:.2f} ms, min {min(times) * 1e3:.2f} ms"}
            if not args.no_configs:
                import bench_configs
                try:
                    line["configs"] = bench_configs.run_all(eng, peak)
                except Exception as e:
                    line["configs"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)  # (no teardown of NCCL / CUDA IPC mappings: nothing left to do, and teardown order has blocked before)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config sub-record (N = 1)")
    ap.add_argument("--no-sharded", action="store_true", help="skip BASELINE configs[4] sharded over the ranks")
    ap.add_argument("--extra", action="store_true", help="(kept for compatibility: the configs record is on by default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
