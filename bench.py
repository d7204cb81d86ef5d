#!/usr/bin/env python
"""bench.py -- headline benchmark of the strided map/permute/reduce hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--extra]

Workload (BASELINE.json configs[1]): Float64 4000x4000  `@strided B .= (A .+ A') ./ 2`.
A "step" is one pass of the hot path over one such problem per GPU (N GPUs = N independent problems, i.e. an
outer batch dimension sharded over the ranks: weak scaling, no data-path collective).
Metric: effective GB/s = ALGORITHMIC bytes / time, algorithmic bytes = compulsory traffic = 8 B x (distinct
input elements + output elements) = 256 000 000 B per problem (A and A' alias the same memory).

  value     device-resident throughput, K steps timed with CUDA events on the launching stream, max over ranks
  e2e       same metric through the C ABI with HOST buffers (sb_mapreduce_host): pinned H2D of A + kernel + D2H
            of B inside the timed region
  roofline  the dominant kernel (map_tile) against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the restated reference CPU path (oracle/, all host threads) on the same workload

`--impl reference` times the reference's own CPU path: Julia is not in this image, so this is the C
restatement of Strided.jl's planner + task bisection + blocked loop nest (oracle/strided_ref.c), kind "port".
"""
import argparse
import json
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
        "config": {"workload": "BASELINE configs[1]: Float64 4000x4000 B .= (A .+ A')./2", "algorithmic_bytes": ALG_BYTES,
                   "note": "restated Strided.jl CPU path (oracle/strided_ref.c): Julia is not installed in this image"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": nthreads, "kind": "port",
                         "sample": f"{len(all_times)} full passes of the 4000x4000 problem, mean; min {min(all_times) * 1e3:.2f} ms",
                         "value_1_thread": ALG_BYTES / float(np.mean(t1)) / 1e9},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * ALG_BYTES / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    # Every step: H2D of A (pinned) -> kernel -> D2H of B, all through sb_mapreduce_host.  Two contexts (two streams,
    # two staging pools, two output buffers) are used alternately in stream-ordered mode, so that the D2H of step n
    # overlaps the H2D of step n+1 on the full-duplex PCIe link; every step still moves all of its bytes.
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
    e2e_value = world * ALG_BYTES / (e2e_s / e2e_steps) / 1e9
    # the same call made synchronously, one at a time (latency view)
    engs[0].set_sync(True)
    t0 = time.perf_counter()
    for _ in range(3):
        sb.run_mapreduce(TOKENS, 0, 0, 0.0, (N_MAT, N_MAT), host_views[0], engine=engs[0])
    e2e_sync_value = world * ALG_BYTES / ((time.perf_counter() - t0) / 3) / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    achieved = ALG_BYTES / (ms_per_step * 1e-3) / 1e9  # per GPU: one map_tile launch per step
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
        "config": {"workload": "BASELINE configs[1]: Float64 4000x4000 B .= (A .+ A')./2, one problem per GPU",
                   "algorithmic_bytes_per_gpu": ALG_BYTES, "operand_bytes_per_gpu": 3 * N_MAT * N_MAT * 8,
                   "l2": "working set 256 MB (A 128 MB + B 128 MB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"{world} independent problems (batch dim sharded), no collective",
                   "timing": "K steps captured in one CUDA graph, one replay timed with CUDA events on the launching stream",
                   "vs_baseline_ref": "README.md:120-121 @strided 4 threads 30.355 ms = 8.43 GB/s, hardware not stated",
                   "plan": plan},
        "gpu_launches": launches_timed,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": est["h2d_bytes"] // e2e_steps,
                "d2h_bytes_per_step": est["d2h_bytes"] // e2e_steps, "steps": e2e_steps,
                "api": "sb_mapreduce_host (C ABI, pinned host buffers), 2 contexts alternating, stream-ordered",
                "one_call_at_a_time": e2e_sync_value},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel, "peak_source": peak_src,
                     "basis": "algorithmic bytes 256e6 per launch / CUDA-event time per launch"},
    }
    if world == 1:
        times, nthreads = time_reference(budget_s=12.0)
        cv = ALG_BYTES / float(np.mean(times)) / 1e9
        t1, _ = time_reference(1.5, min_reps=2, nthreads=1)
        line["cpu_baseline"] = {"value": cv, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                "value_1_thread": ALG_BYTES / float(np.mean(t1)) / 1e9,
                                "sample": f"{len(times)} full passes of the same 4000x4000 problem (restated Strided.jl CPU path, "
                                          f"{nthreads} tasks), mean {np.mean(times) * 1e3:.2f} ms, min {min(times) * 1e3:.2f} ms"}
    if args.extra:
        import bench_configs
        extra = bench_configs.run_all(eng, peak)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(extra, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)
        print(json.dumps(extra), file=sys.stderr)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--extra", action="store_true", help="also measure the other BASELINE configs -> gpurun_out/bench_configs.json")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
