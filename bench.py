#!/usr/bin/env python3
"""Benchmark of the batched RMHMC hot path (BASELINE.json: min-ESS/sec & leapfrog steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload german|australian] [--chains C] [--rounds-per-step R]

One STEP = R rounds; one round = one generalized leapfrog step (rmhmc.py:96-163) for each of the C
chains of a rank, plus the accept/reject and the next momentum draw of every chain whose
trajectory ends in that round.  Chains free-run (no lock-step over MCMC iterations).

Workload (default): German-credit-shaped synthetic logistic regression, N=1000, D=25
(BASELINE.json configs[0]/[3]; the north_star's target is quoted on it), RMHMC with
NumOfLeapFrogSteps=6, StepSize=0.5, NumOfNewtonSteps=6, 65536 chains PER GPU (weak scaling).

Printed JSON line (rank 0): value = min over parameters of the ESS summed over all chains of all
ranks (tools.CalculateESS semantics per chain, on the samples drawn inside the timed region)
divided by the timed seconds (max over ranks).  ESS post-processing is outside the timed region,
as in the reference (main.py:70-79 runs after the samplers' own timers).

--impl reference times the CPU arm: the oracle port of rmhmc.py (bit-identical to the reference in
the build container; /root/reference itself is not available on the GPU box) on all host cores,
one independent chain per core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape key, description, default chains per GPU, default rounds per step)
    "german": ("german", "German-credit-shaped synthetic logistic regression N=1000 D=25 (seed 1234), "
               "RMHMC L=6 eps=0.5 F=6", 65536, 50),
    "australian": ("australian", "Australian-credit-shaped synthetic logistic regression N=690 D=15 (seed 1235), "
                   "RMHMC L=6 eps=0.5 F=6", 4096, 400),
    # BASELINE.json configs[2]; a round takes ~0.46 s, so a step is 2 rounds (use --steps 8); the CPU arm is omitted
    # (one reference iteration at this size takes ~20 s)
    "cfg3": ("synthetic:100000:100:1236", "synthetic logistic regression N=100000 D=100 (seed 1236), RMHMC L=6 eps=0.5 F=6",
             1024, 2),
}
N_LEAPFROG, STEP_SIZE, N_FIXED = 6, 0.5, 6


def load_data(key):
    """(XX, t) of a workload: a named shape or ``synthetic:N:D:seed`` (SURVEY.md 8d generator)."""
    from riemannhamiltonianmontecarlo_b200 import datasets
    if key.startswith("synthetic:"):
        _, n, d, seed = key.split(":")
        return datasets.synthetic_logistic(int(n), int(d), int(seed))
    return datasets.shaped(key)
METRIC = "min_ess_per_sec"
UNIT = "ESS/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="german", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: per workload)")
    ap.add_argument("--rounds-per-step", type=int, default=0)
    ap.add_argument("--partials", default="matrix_free", choices=["matrix_free", "tensor"],
                    help="how the engine evaluates the metric partials (include/rmhmc_b200.h)")
    ap.add_argument("--metric", default=None, choices=["dmma", "i8"],
                    help="metric build: FP64 DMMA kernel or INT8-slice tcgen05 build (default: the library's)")
    ap.add_argument("--ref-iters-per-step", type=int, default=30)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-iters", type=int, default=450)
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU arm
def _ref_worker(args):
    """One core: warm up, then time `steps * ips` iterations of the oracle port of rmhmc.py."""
    idx, shape, warm_iters, timed_iters, barrier_path = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import blr_oracle as bo
    xx, t = load_data(shape)
    d = xx.shape[1]
    tape_w = bo.make_tape(max(warm_iters, 1), d, 50_000 + idx)
    _, info = bo.rmhmc_chain(xx, t, tape_w, n_iter=max(warm_iters, 1), burn_in=0, n_leapfrog=N_LEAPFROG,
                             step_size=STEP_SIZE, n_fixed=N_FIXED)
    tape = bo.make_tape(timed_iters, d, 60_000 + idx)
    t0 = time.perf_counter()
    samples, info2 = bo.rmhmc_chain(xx, t, tape, n_iter=timed_iters, burn_in=0, n_leapfrog=N_LEAPFROG,
                                    step_size=STEP_SIZE, n_fixed=N_FIXED, w0=info["w"])
    dt = time.perf_counter() - t0
    s = samples[1:]
    ess = bo.ess(s, s.shape[0] - 1)[:, 0]
    return {"seconds": dt, "ess": ess.tolist(), "leapfrogs": int(info2["steps"].sum()),
            "iters": int(timed_iters), "accept": float(info2["accepted"].mean())}


def run_cpu_arm(shape, steps, warmup, iters_per_step, cores):
    """The reference's algorithm on `cores` host cores, one independent chain each."""
    import multiprocessing as mp
    import numpy as np
    warm_iters, timed_iters = warmup * iters_per_step, steps * iters_per_step
    jobs = [(i, shape, warm_iters, timed_iters, None) for i in range(cores)]
    if cores == 1:
        res = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_ref_worker, jobs)
    seconds = max(r["seconds"] for r in res)
    ess_sum = np.sum([r["ess"] for r in res], axis=0)
    return {
        "seconds": seconds,
        "min_ess_per_sec": float(ess_sum.min() / seconds),
        "leapfrog_per_sec": float(sum(r["leapfrogs"] for r in res) / seconds),
        "iters_per_sec": float(sum(r["iters"] for r in res) / seconds),
        "accept": float(np.mean([r["accept"] for r in res])),
        "cores": cores,
        "sample": f"{cores} chain(s) x {timed_iters} iterations after {warm_iters} warm-up iterations, "
                  f"oracle port of rmhmc.py (numpy/OpenBLAS, 1 BLAS thread per chain)",
    }


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape, descr, _, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    out = run_cpu_arm(shape, args.steps, args.warmup, args.ref_iters_per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": out["min_ess_per_sec"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": out["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": descr, "chains": cores, "iterations_per_step": args.ref_iters_per_step},
        "leapfrog_steps_per_sec": out["leapfrog_per_sec"], "iterations_per_sec": out["iters_per_sec"],
        "accept_rate": out["accept"],
        "cpu_baseline": {"value": out["min_ess_per_sec"], "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": out["sample"]},
        "e2e": {"value": out["min_ess_per_sec"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ GPU arm
def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM burst on this GPU (TFLOP/s): the FP64 tensor-pipe denominator, measured live."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = float("inf")
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def gpu_main(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import riemannhamiltonianmontecarlo_b200 as r
    from riemannhamiltonianmontecarlo_b200.engine import ess_ragged

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    shape, descr, c_default, r_default = WORKLOADS[args.workload]
    C = args.chains or c_default
    R = args.rounds_per_step or r_default
    K, W = args.steps, args.warmup
    xx, t = load_data(shape)
    N, D = xx.shape
    P2, P3 = D * (D + 1) // 2, D * (D + 1) * (D + 2) // 6

    # pinned host copies: the e2e arm uploads them every step
    xx_host = torch.from_numpy(xx).pin_memory()
    t_host = torch.from_numpy(t.reshape(-1)).pin_memory()
    xx_dev = torch.empty_like(xx_host, device=device)
    t_dev = torch.empty_like(t_host, device=device)

    data = r.LogisticData(xx, t, device=device, partials=args.partials, metric=args.metric)
    sampler = r.RMHMCSampler(data, C, N_LEAPFROG, STEP_SIZE, N_FIXED)
    sampler.set_philox(20261018, chain_offset=rank * C)
    # the sample store holds every iteration of the run: bound it (ESS kernel: <= 24000 rows; HBM) by shortening the
    # step when many steps are requested
    passes = W + K * (1 if args.no_e2e else 2)
    max_rows = 6000
    if passes * R / 3.5 * 1.12 + 96 > max_rows:
        R = max(1, int((max_rows - 96) * 3.5 / 1.12 / passes))
    total_rounds = passes * R
    cap = int(total_rounds / 3.5 * 1.12) + 96           # E[RandomStep] = 3.5 rounds per iteration
    samples = sampler.set_samples(cap, 0)                # row it = state after iteration it

    def iters_now():
        st_i = torch.empty(C, dtype=torch.int64, device=device)
        st_l = torch.empty(C, dtype=torch.int64, device=device)
        from ctypes import c_void_p
        from riemannhamiltonianmontecarlo_b200 import _capi
        _capi.check(sampler._lib.rmhmc_read_state(sampler.h, c_void_p(0), c_void_p(st_i.data_ptr()), c_void_p(0),
                                                  c_void_p(st_l.data_ptr()), c_void_p(0), c_void_p(0)), sampler.h, "read_state")
        return st_i, st_l

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def reduce_max(x):
        tt = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def reduce_sum(tt):
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return tt

    def window_stats(it0, it1):
        """min_d sum_c ESS and R-hat over the rows each chain produced in [it0, it1)."""
        counts = torch.clamp(torch.minimum(it1, torch.full_like(it1, cap)) - it0, min=0)
        ess = ess_ragged(samples, it0, counts)                       # (C, D) on the GPU
        ess = torch.nan_to_num(ess, nan=0.0)       # a chain frozen over the whole window (0/0 in tools.py:27) counts as 0
        ess_sum = reduce_sum(ess.sum(dim=0))
        # Gelman-Rubin over the rows every chain has in common (cheap summary, torch plumbing)
        lo, hi = int(it0.max().item()), int(torch.minimum(it1, torch.full_like(it1, cap)).min().item())
        rhat_max = None
        if hi - lo >= 4:
            win = samples[:, lo:hi, :]
            n = hi - lo
            means = win.mean(dim=1)
            varis = win.var(dim=1, unbiased=True)
            stat = torch.stack([means.sum(0), (means ** 2).sum(0), varis.sum(0)])
            stat = reduce_sum(stat)
            m_tot = C * world
            w_ = stat[2] / m_tot
            b_over_n = (stat[1] - stat[0] ** 2 / m_tot) / (m_tot - 1)
            rhat_max = float(torch.sqrt(((n - 1) / n * w_ + b_over_n) / w_).max().item())
        return ess_sum, rhat_max, int(counts.sum().item())

    # ---------------------------------------------------------------- warm-up (also the burn-in)
    for _ in range(max(W, 0)):
        sampler.advance(R)
    barrier()
    fp64_peak = measure_fp64_peak(torch, device)

    # ---------------------------------------------------------------- timed region: inputs resident in HBM
    it0, lf0 = iters_now()
    launches0 = sampler.launch_count()
    sampler.profile(True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sampler.advance(R)
    e1.record()
    barrier()
    seconds = reduce_max(e0.elapsed_time(e1) * 1e-3)
    clock_info = clocks.stop() if rank == 0 else None
    prof = sampler.profile_read()
    sampler.profile(False)
    launches = sampler.launch_count() - launches0
    it1, lf1 = iters_now()
    ess_sum, rhat_max, n_samples = window_stats(it0, it1)
    leapfrogs = float(reduce_sum((lf1 - lf0).sum().to(torch.float64)).item())
    iters_done = float(reduce_sum((it1 - it0).sum().to(torch.float64)).item())
    launches_all = float(reduce_sum(torch.tensor(float(launches), dtype=torch.float64, device=device)).item())
    value = float(ess_sum.min().item()) / seconds

    # ---------------------------------------------------------------- e2e: host buffers in, host samples out, every step
    e2e = None
    if not args.no_e2e:
        rows_per_step = int(R / 3.5 * 1.5) + 64
        host_out = torch.empty(C * rows_per_step * D, dtype=torch.float64).pin_memory()
        dev_stage = torch.empty(C * rows_per_step * D, dtype=torch.float64, device=device)
        it_a, _ = iters_now()
        it_prev = it_a
        d2h_bytes = 0
        barrier()
        t_start = time.perf_counter()
        for _ in range(K):
            xx_dev.copy_(xx_host, non_blocking=True)                  # H2D of the step's inputs
            t_dev.copy_(t_host, non_blocking=True)
            data.update(xx_dev, t_dev)
            sampler.advance(R)
            it_now, _ = iters_now()                                   # synchronises
            lo = int(it_prev.min().item())
            hi = min(int(it_now.max().item()), cap)
            n_rows = max(min(hi - lo, rows_per_step), 0)
            # D2H of the step's samples: pack the (strided) rows on the device, then ONE contiguous copy into pinned memory
            n_el = C * n_rows * D
            dev_stage[:n_el].view(C, n_rows, D).copy_(samples[:, lo:lo + n_rows])
            host_out[:n_el].copy_(dev_stage[:n_el], non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            d2h_bytes += n_el * 8
            it_prev = it_now
        torch.cuda.synchronize(device)
        e2e_seconds = reduce_max(time.perf_counter() - t_start)
        barrier()
        ess_sum_e, _, _ = window_stats(it_a, it_prev)
        e2e = {"value": float(ess_sum_e.min().item()) / e2e_seconds, "unit": UNIT,
               "h2d_bytes_per_step": int(xx_host.numel() * 8 + t_host.numel() * 8),
               "d2h_bytes_per_step": int(d2h_bytes / K), "seconds": e2e_seconds}

    st = sampler.state()
    accept = float(st["accepted"].sum() / max(st["iters"].sum(), 1))
    renorm = int(st["renorm_momentum"].sum() + st["renorm_position"].sum())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------------------------------------------------------- roofline (dominant kernel of the mode)
    # algorithmic flops per launch (packed symmetric contractions)
    alg_flops = {"metric_fp": 2.0 * C * N * (P2 + D), "metric_closing": 2.0 * C * N * (P2 + 2 * D),
                 "partials": 2.0 * C * N * P3, "quad_pass": 4.0 * C * N * D, "leverage_gemm": 2.0 * C * N * P2,
                 "trace_pass": 2.0 * C * N * D}
    if args.partials == "tensor":
        top, top_name = "partials", "k_tbuild_pre (partials build T = Cw . KR3(X), FP64 DMMA.8x8x4)"
        w_alg = 2.0 * N * P3 + 2.0 * N_FIXED * N * P2         # SURVEY.md 8d, per chain-leapfrog-step
    else:
        top, top_name = "metric_fp", "k_metric<MODE 0> (f = X theta, G = V . KR2(X), FP64 DMMA.8x8x4)"
        # F metric builds + leverage GEMM + (F + 1) quadratic-form passes + trace pass
        w_alg = 2.0 * N_FIXED * N * (P2 + D) + 2.0 * N * P2 + (N_FIXED + 1) * 4.0 * N * D + 2.0 * N * D
    # DRAM traffic of the dominant kernel per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    # `ncu --set full` captures of exactly these configurations; null for any other configuration
    NCU_TRAFFIC = {
        ("german", 65536, "matrix_free"): (13.44e6 + 117.75e6, "profiles/r01/ncu_v3_metric_fp_raw.csv"),
        ("german", 65536, "tensor"): (539.19e6 + 1487.65e6, "profiles/r01/ncu_v2_tbuild_raw.csv"),
    }
    traffic, traffic_src = NCU_TRAFFIC.get((args.workload, C, args.partials), (None, None))
    top_ms, top_n = prof[top]
    flops_per_launch = alg_flops[top]
    achieved = flops_per_launch / (top_ms / max(top_n, 1) * 1e-3) / 1e12 if top_n else None
    microbench_peak = 37.1                                     # profiles/microbench/r01_fp64_peak_b200.txt
    peak = max(fp64_peak, microbench_peak)
    kernels = {}
    wall_ms = seconds * 1e3
    for name, (ms, n) in prof.items():
        if not n:
            continue
        kernels[name] = {"launches": n, "ms_total": ms, "ms_avg": ms / max(n, 1), "share_of_step": ms / wall_ms}
        if name in alg_flops and n:
            kernels[name]["tflops_alg"] = alg_flops[name] / (ms / n * 1e-3) / 1e12
    NP = (N + 31) // 32 * 32
    if args.partials == "tensor":
        ws_note = "per-round working set (T slots + cbuf, %.1f GB) exceeds the 126 MB L2" % (
            (2 * C * (P3 + 8) * 8 + C * NP * 8) / 1e9)
    else:
        ws_note = "per-round working set (c_n slots + leverages + G^-1/L per chain, %.1f GB) exceeds the 126 MB L2" % (
            (3 * C * NP * 8 + 4 * C * D * D * 8) / 1e9)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": seconds / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": descr, "chains_per_gpu": C, "chains_total": C * world, "rounds_per_step": R,
                   "rng": "philox4x32-10 on device", "partials": args.partials, "l2": ws_note,
                   "baseline_config": {"german": "BASELINE.json configs[3] at this GPU count (the shape the north_star target is "
                                                 "quoted on; configs[0] is the same shape with 1 chain on the CPU)",
                                       "australian": "BASELINE.json configs[1]", "cfg3": "BASELINE.json configs[2]"}[args.workload]},
        "leapfrog_steps_per_sec": leapfrogs / seconds,
        "iterations_per_sec": iters_done / seconds,
        "samples_in_timed_region": n_samples, "accept_rate": accept, "renorm_events": renorm,
        "rhat_max": rhat_max,
        "alg_tflops_overall": w_alg * leapfrogs / seconds / 1e12 / world,
        "alg_flops_per_chain_leapfrog": w_alg,
        "roofline": {"bound": "tensor", "kernel": top_name,
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "max(cuBLAS DGEMM 4096^3 measured live = %.1f, DMMA issue microbenchmark = %.1f); "
                                    "MEASURED_PEAKS.json has no FP64 entry" % (fp64_peak, microbench_peak),
                     "flops_per_launch": flops_per_launch},
        "kernels": kernels,
        "gpu_launches": int(launches_all),
        "clocks": clock_info,
        "e2e": e2e,
    }
    if args.workload == "cfg3":
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                "sample": "omitted: one iteration of the reference at N=1e5, D=100 takes ~20 s"}
    elif not args.no_cpu_baseline and world == 1:
        cb = run_cpu_arm(shape, 1, 1, args.cpu_baseline_iters // 2, 1)
        line["cpu_baseline"] = {"value": cb["min_ess_per_sec"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": cb["sample"], "leapfrog_steps_per_sec": cb["leapfrog_per_sec"],
                                "iterations_per_sec": cb["iters_per_sec"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_main(args)
    return gpu_main(args)


if __name__ == "__main__":
    sys.exit(main())
