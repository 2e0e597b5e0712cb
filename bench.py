#!/usr/bin/env python3
"""Benchmark of the batched RMHMC hot path (BASELINE.json: min-ESS/sec & leapfrog steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload german|australian|cfg3] [--sampler rmhmc|hmc|mmala|mmala_simp]
                    [--chains C] [--scaling strong|weak] [--metric i8|dmma] [--partials matrix_free|tensor]

One STEP = R rounds.  RMHMC: one round = one generalized leapfrog step (rmhmc.py:96-163) for each chain of the rank,
plus the accept/reject and the next momentum draw of every chain whose trajectory ends in that round; chains free-run
(no lock-step over MCMC iterations).  HMC: one round = one leapfrog step (hmc.py:51-62); mMALA: one iteration.

Workload (default): German-credit-shaped synthetic logistic regression, N=1000, D=25, RMHMC with
NumOfLeapFrogSteps=6, StepSize=0.5, NumOfNewtonSteps=6, 65 536 chains IN TOTAL partitioned over the N GPUs
(BASELINE.json configs[3]: "65536 chains sharded over 1/2/4/8 B200" -> "scaling": "strong"; --scaling weak keeps
65 536 chains per GPU).  There is no collective on the data path: NCCL only combines the end-of-run statistics
(rmhmc_stats_gather inside the library).

Printed JSON line (rank 0): value = min over parameters of the ESS summed over all chains of all ranks
(tools.CalculateESS semantics per chain, on the samples drawn inside the timed region) divided by the timed seconds
(max over ranks).  ESS post-processing is outside the timed region, as in the reference (main.py:70-79 runs after the
samplers' own timers).

--impl reference times the CPU arm: the oracle port of the sampler (bit-identical to the reference in the build
container; /root/reference itself is not available on the GPU box) on all host cores, one independent chain per core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape key, description, default chains IN TOTAL, default rounds per step)
    "german": ("german", "German-credit-shaped synthetic logistic regression N=1000 D=25 (seed 1234)", 65536, 50),
    "australian": ("australian", "Australian-credit-shaped synthetic logistic regression N=690 D=15 (seed 1235)", 4096, 400),
    # BASELINE.json configs[2]; a round takes ~0.25 s, so a step is 2 rounds (use --steps 8); the CPU arm is omitted
    # (one reference iteration at this size takes ~20 s)
    "cfg3": ("synthetic:100000:100:1236", "synthetic logistic regression N=100000 D=100 (seed 1236)", 1024, 2),
}
N_LEAPFROG, STEP_SIZE, N_FIXED = 6, 0.5, 6
# HMC: the MATLAB originals' per-dataset step sizes (BLR_hmc.m:36,72); hmc.py's default 0.14 gives 0 % acceptance here
HMC_STEP = {"german": 0.05, "australian": 0.1, "cfg3": 0.01}
HMC_LEAPFROG = 100
MMALA_STEP = 1.0
# expected rounds per MCMC iteration (RandomStep is uniform on 1..L)
ROUNDS_PER_ITER = {"rmhmc": 3.5, "hmc": 50.5, "mmala": 1.0, "mmala_simp": 1.0}
SAMPLER_DESCR = {"rmhmc": "RMHMC L=6 eps=0.5 F=6", "hmc": "HMC L=100 eps=%g (BLR_hmc.m per-dataset step)",
                 "mmala": "mMALA eps=1", "mmala_simp": "simplified mMALA eps=1"}
METRIC = "min_ess_per_sec"
UNIT = "ESS/s"


def load_data(key):
    """(XX, t) of a workload: a named shape or ``synthetic:N:D:seed`` (SURVEY.md 8d generator)."""
    from riemannhamiltonianmontecarlo_b200 import datasets
    if key.startswith("synthetic:"):
        _, n, d, seed = key.split(":")
        return datasets.synthetic_logistic(int(n), int(d), int(seed))
    return datasets.shaped(key)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="german", choices=sorted(WORKLOADS))
    ap.add_argument("--sampler", default="rmhmc", choices=sorted(ROUNDS_PER_ITER))
    ap.add_argument("--chains", type=int, default=0, help="chains IN TOTAL (strong) / per GPU (weak); default per workload")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the chains are partitioned over the GPUs (BASELINE.json configs[3]); weak: per GPU")
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
def _oracle_chain(bo, sampler, xx, t, tape, n_iter, w0, workload):
    if sampler == "rmhmc":
        return bo.rmhmc_chain(xx, t, tape, n_iter=n_iter, burn_in=0, n_leapfrog=N_LEAPFROG, step_size=STEP_SIZE,
                              n_fixed=N_FIXED, w0=w0)
    if sampler == "hmc":
        return bo.hmc_chain(xx, t, tape, n_iter=n_iter, burn_in=0, n_leapfrog=HMC_LEAPFROG, step_size=HMC_STEP[workload], w0=w0)
    return bo.mmala_chain(xx, t, tape, n_iter=n_iter, burn_in=0, step_size=MMALA_STEP, simplified=sampler == "mmala_simp", w0=w0)


def _ref_worker(args):
    """One core: warm up, then time `steps * ips` iterations of the oracle port of the sampler."""
    idx, workload, sampler, warm_iters, timed_iters = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import blr_oracle as bo
    xx, t = load_data(WORKLOADS[workload][0])
    d = xx.shape[1]
    tape_w = bo.make_tape(max(warm_iters, 1), d, 50_000 + idx)
    _, info = _oracle_chain(bo, sampler, xx, t, tape_w, max(warm_iters, 1), None, workload)
    tape = bo.make_tape(timed_iters, d, 60_000 + idx)
    t0 = time.perf_counter()
    samples, info2 = _oracle_chain(bo, sampler, xx, t, tape, timed_iters, info["w"], workload)
    dt = time.perf_counter() - t0
    s = samples[1:]
    ess = bo.ess(s, s.shape[0] - 1)[:, 0]
    steps = info2["steps"].sum() if "steps" in info2 else timed_iters
    return {"seconds": dt, "ess": ess.tolist(), "leapfrogs": int(steps),
            "iters": int(timed_iters), "accept": float(info2["accepted"].mean())}


def run_cpu_arm(workload, sampler, steps, warmup, iters_per_step, cores):
    """The reference's algorithm on `cores` host cores, one independent chain each."""
    import multiprocessing as mp
    import numpy as np
    warm_iters, timed_iters = warmup * iters_per_step, steps * iters_per_step
    jobs = [(i, workload, sampler, warm_iters, timed_iters) for i in range(cores)]
    if cores == 1:
        res = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_ref_worker, jobs)
    seconds = max(r["seconds"] for r in res)
    ess_sum = np.nansum([r["ess"] for r in res], axis=0)
    port = {"rmhmc": "rmhmc.py", "hmc": "hmc.py"}.get(sampler, "BLR_mMALA.m (parity unpinned: MATLAB only)")
    return {
        "seconds": seconds,
        "min_ess_per_sec": float(ess_sum.min() / seconds),
        "leapfrog_per_sec": float(sum(r["leapfrogs"] for r in res) / seconds),
        "iters_per_sec": float(sum(r["iters"] for r in res) / seconds),
        "accept": float(np.mean([r["accept"] for r in res])),
        "cores": cores,
        "sample": f"{cores} chain(s) x {timed_iters} iterations after {warm_iters} warm-up iterations, "
                  f"oracle port of {port} (numpy/OpenBLAS, 1 BLAS thread per chain)",
    }


def sampler_descr(args):
    d = SAMPLER_DESCR[args.sampler]
    return d % HMC_STEP[args.workload] if args.sampler == "hmc" else d


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape, descr, _, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    ips = args.ref_iters_per_step
    if args.sampler != "rmhmc":
        ips = max(ips, 60)
    out = run_cpu_arm(args.workload, args.sampler, args.steps, args.warmup, ips, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": out["min_ess_per_sec"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": out["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": descr + ", " + sampler_descr(args), "sampler": args.sampler, "chains": cores,
                   "iterations_per_step": ips},
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
def make_sampler(r, args, data, chains, rank_offset):
    if args.sampler == "rmhmc":
        s = r.RMHMCSampler(data, chains, N_LEAPFROG, STEP_SIZE, N_FIXED)
    elif args.sampler == "hmc":
        s = r.HMCSampler(data, chains, HMC_LEAPFROG, HMC_STEP[args.workload])
    else:
        s = r.MMALASampler(data, chains, MMALA_STEP, args.sampler == "mmala_simp")
    s.set_philox(20261018, chain_offset=rank_offset)
    return s


def gpu_main(args):
    import torch
    import torch.distributed as dist

    import riemannhamiltonianmontecarlo_b200 as r
    from riemannhamiltonianmontecarlo_b200.engine import device_peaks, ess_ragged, shard_chains

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
    c_arg = args.chains or c_default
    if args.scaling == "strong":
        c_total = c_arg
        c0, c1 = shard_chains(c_total, rank, world)          # this rank's contiguous chain range
    else:
        c_total = c_arg * world
        c0, c1 = rank * c_arg, (rank + 1) * c_arg
    C = c1 - c0
    rpi = ROUNDS_PER_ITER[args.sampler]
    R = args.rounds_per_step or {"rmhmc": r_default, "hmc": 8 * r_default, "mmala": max(r_default // 2, 1),
                                 "mmala_simp": max(r_default // 2, 1)}[args.sampler]
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
    metric_mode = data.metric_mode
    data.init_stats_comm(rank, world)                        # NCCL communicator inside the library: statistics only
    sampler = make_sampler(r, args, data, C, c0)
    # the sample store holds every iteration of the run: bound it (ESS kernel: <= 24000 rows; HBM) by shortening the
    # step when many steps are requested
    passes = W + K * (1 if args.no_e2e else 2)
    max_rows = 6000
    if passes * R / rpi * 1.12 + 96 > max_rows:
        R = max(1, int((max_rows - 96) * rpi / 1.12 / passes))
    total_rounds = passes * R
    cap = int(total_rounds / rpi * 1.12) + 96
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

    def window_stats(it0, it1, extra):
        """min_d sum_c ESS, R-hat and the summed scalars over the rows each chain produced in [it0, it1): per-chain ESS by
        the library's ragged ESS kernel, everything combined over the ranks by rmhmc_stats_gather (NCCL in the library)."""
        counts = torch.clamp(torch.minimum(it1, torch.full_like(it1, cap)) - it0, min=0)
        ess = ess_ragged(samples, it0, counts)                       # (C, D) on the GPU; NaN = frozen chain
        # Gelman-Rubin over the rows every chain of every rank has in common
        lo = int(reduce_max(float(it0.max().item())))
        hi = -int(reduce_max(-float(torch.minimum(it1, torch.full_like(it1, cap)).min().item())))
        win = samples[:, lo:hi, :] if hi - lo >= 4 else None
        scal = torch.tensor([float(counts.sum().item())] + list(extra), dtype=torch.float64, device=device)
        ess_sum, rhat, scal = data.stats_gather(ess=ess, samples=win, scalars=scal)
        rhat_max = float(rhat.max().item()) if rhat is not None else None
        return ess_sum, rhat_max, scal.cpu().numpy()

    # ---------------------------------------------------------------- warm-up (also the burn-in)
    for _ in range(max(W, 0)):
        sampler.advance(R)
    barrier()
    dmma_peak, i8_peak = device_peaks(device)

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
    ess_sum, rhat_max, scal = window_stats(it0, it1, [float((lf1 - lf0).sum().item()), float((it1 - it0).sum().item()),
                                                      float(launches)])
    n_samples, leapfrogs, iters_done, launches_all = int(scal[0]), float(scal[1]), float(scal[2]), float(scal[3])
    value = float(ess_sum.min().item()) / seconds

    # ---------------------------------------------------------------- e2e: host buffers in, host samples out, every step
    e2e = None
    if not args.no_e2e:
        rows_per_step = int(R / rpi * 1.5) + 64
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
        ess_sum_e, _, _ = window_stats(it_a, it_prev, [])
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

    # ---------------------------------------------------------------- rooflines (per kernel class)
    # algorithmic work per launch: packed symmetric contractions in FP64 flops; the INT8 GEMM in digit multiply-adds
    # (15 = S (S + 1) / 2 digit products for S = 5 digits per operand)
    F = N_FIXED
    n_digit_products = 15 if os.environ.get("RMHMC_I8_SLICES", "5") != "6" else 21
    rm = args.sampler == "rmhmc"
    work = {
        "metric_fp": (2.0 * C * N * (P2 + D), "fp64"), "metric_closing": (2.0 * C * N * (P2 + 2 * D), "fp64"),
        "partials": (2.0 * C * N * P3, "fp64"), "quad_pass": (4.0 * C * N * D * (F if rm else 1), "fp64"),
        "leverage_gemm": (2.0 * C * N * P2, "fp64"), "trace_pass": (2.0 * C * N * D * (3 if rm else 1), "fp64"),
        "i8_gemm": (2.0 * n_digit_products * C * N * P2, "i8"),
        "i8_vslice": (2.0 * C * N * (D + 20), "fp64"),      # f = X theta plus ~20 FP64 operations of the logistic terms and digits
        # closing build: the same plus X^T (t - p) and ~14 FP64 operations of t - p, log(1 + e), c_n
        "i8_vslice_closing": (2.0 * C * N * (2 * D + 34), "fp64"),
    }
    if args.sampler == "hmc":
        work["metric_closing"] = (4.0 * C * N * D, "fp64")          # k_metric<MODE 2>: f = X theta and X^T (t - p)
        if os.environ.get("RMHMC_HMC_FUSED", "1") != "0" and D <= 32:
            # k_hmc_rounds: the same two contractions for every round of the launch (64 rounds per launch)
            work["chain_turn"] = (4.0 * C * N * D * R / ((R + 63) // 64), "fp64")
    if metric_mode == "i8" and args.sampler != "hmc":        # kinds 0 / 1 are the sums of the two i8 kernels there
        work.pop("metric_fp"); work.pop("metric_closing")
        if args.partials == "matrix_free" and os.environ.get("RMHMC_I8_LEVERAGE", "1") != "0":
            work["leverage_gemm"] = (2.0 * n_digit_products * C * N * P2, "i8")     # the same digit GEMM, K = packed pairs
    names = {
        "metric_fp": "k_metric<MODE 0> (f = X theta, G = V . KR2(X), FP64 DMMA.8x8x4)",
        "metric_closing": "k_metric<MODE 1|2> (closing build / HMC gradient, FP64 DMMA.8x8x4)",
        "partials": "k_tbuild_pre (partials build T = Cw . KR3(X), FP64 DMMA.8x8x4)",
        "quad_pass": "k_pass<MOMFP> (implicit momentum fixed point: F quadratic-form passes, FP64 DMMA.8x8x4)",
        "leverage_gemm": "leverage GEMM h = q . KR2(X)^T (k_i8_gemm in INT8 metric mode, else k_tbuild_pre on FP64 DMMA.8x8x4)",
        "trace_pass": "k_pass<PAIR|TRACE> (tr(G^-1 dG_d) and u^T dG_d u passes, FP64 DMMA.8x8x4)",
        "i8_gemm": "k_i8_gemm (G = V . KR2(X) as 15 exact INT8 digit GEMMs: tcgen05.mma.kind::i8, TMEM, tensor-map TMA)",
        "i8_vslice": "k_i8_vslice_mma (position iterates: f = X theta on DMMA.8x8x4, logistic terms, base-256 digits of v; FP64)",
        "i8_vslice_closing": "k_i8_vslice_mma_closing (closing build: f, digits of v, X^T (t - p), log-likelihood, c_n; FP64)",
        "chain_turn": "k_hmc_rounds (64 leapfrog rounds per launch: f = X w, X^T (t - sigma(f)) on FP64 DMMA.8x8x4, chain state in registers)",
    }
    peaks = {"fp64": (dmma_peak, "TFLOP/s", "DMMA.8x8x4 issue peak measured live (blr_device_peaks)"),
             "i8": (i8_peak, "TOP/s", "tcgen05.mma.kind::i8 128x256x32 issue peak measured live (blr_device_peaks)")}
    kernels = {}
    wall_ms = seconds * 1e3
    top, top_share = None, -1.0
    for name, (ms, n) in prof.items():
        if not n:
            continue
        kernels[name] = {"launches": n, "ms_total": ms, "ms_avg": ms / max(n, 1), "share_of_step": ms / wall_ms}
        if name in work:
            w_, kind = work[name]
            rate = w_ / (ms / n * 1e-3) / 1e12
            kernels[name].update({"achieved": rate, "unit": peaks[kind][1],
                                  "frac_of_peak": rate / peaks[kind][0] if peaks[kind][0] else None})
            if ms / wall_ms > top_share:
                top, top_share = name, ms / wall_ms
    if args.partials == "tensor":
        w_alg = 2.0 * N * P3 + 2.0 * F * N * P2         # SURVEY.md 8d, per chain-leapfrog-step
    else:
        # F metric builds + leverage GEMM + (F + 1) quadratic-form passes + trace pass
        w_alg = 2.0 * F * N * (P2 + D) + 2.0 * N * P2 + (F + 1) * 4.0 * N * D + 2.0 * N * D
    # DRAM traffic of the kernels per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    # `ncu --set full` captures of exactly this configuration; null for any other configuration
    NCU_TRAFFIC = {
        ("german", 65536, "metric_fp"): (13.44e6 + 117.75e6, "profiles/r01/ncu_v3_metric_fp_raw.csv"),
        ("german", 65536, "partials"): (539.19e6 + 1487.65e6, "profiles/r01/ncu_v2_tbuild_raw.csv"),
        ("german", 65536, "i8_gemm"): (337.59e6 + 144.97e6, "profiles/r02/ncu_r02_i8_gemm_metric_raw.csv"),
        ("german", 65536, "quad_pass"): (5578.80e6 + 38.33e6, "profiles/r02/ncu_r02_final2_mom_fixed_point_raw.csv"),
        ("german", 65536, "i8_vslice"): (25.08e6 + 286.99e6, "profiles/r02/ncu_r02_i8_vslice_iterate_raw.csv"),
        ("german", 65536, "chain_solve"): (212.91e6 + 11.90e6, "profiles/r02/ncu_r02_chain_solve_raw.csv"),
    }
    traffic, traffic_src = NCU_TRAFFIC.get((args.workload, C, top), (None, None))
    roof = None
    if top:
        kind = work[top][1]
        roof = {"bound": "tensor", "kernel": names[top], "achieved": kernels[top]["achieved"], "peak": peaks[kind][0],
                "unit": peaks[kind][1], "frac": kernels[top]["frac_of_peak"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks[kind][2] + "; MEASURED_PEAKS.json has no FP64 / INT8 entry",
                "work_per_launch": work[top][0], "share_of_step": top_share,
                "selection": "the kernel class with the largest share of the timed region; every class is listed under `kernels`"}
    NP = (N + 31) // 32 * 32
    ws_note = "per-round working set (c_n slots + leverages + G^-1/L per chain, %.1f GB) %s the 126 MB L2" % (
        (3 * C * NP * 8 + 4 * C * D * D * 8) / 1e9, "exceeds" if 3 * C * NP * 8 > 126e6 else "fits")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": seconds / K * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": descr + ", " + sampler_descr(args), "sampler": args.sampler, "chains_per_gpu": C,
                   "chains_total": c_total, "rounds_per_step": R, "rng": "philox4x32-10 on device", "partials": args.partials,
                   "metric_build": {"i8": "INT8 digits on tcgen05 (FP64 result, 5 balanced base-256 digits per operand)",
                                    "dmma": "FP64 DMMA"}[metric_mode], "l2": ws_note,
                   "baseline_config": {"german": "BASELINE.json configs[3] at this GPU count (the shape the north_star target is "
                                                 "quoted on; configs[0] is the same shape with 1 chain on the CPU)",
                                       "australian": "BASELINE.json configs[1]", "cfg3": "BASELINE.json configs[2]"}[args.workload]},
        "leapfrog_steps_per_sec": leapfrogs / seconds,
        "iterations_per_sec": iters_done / seconds,
        "samples_in_timed_region": n_samples, "accept_rate": accept, "renorm_events": renorm,
        "rhat_max": rhat_max,
        "alg_tflops_overall": (w_alg * leapfrogs / seconds / 1e12 / world) if rm else None,
        "alg_flops_per_chain_leapfrog": w_alg if rm else 4.0 * N * D,
        "roofline": roof,
        "peaks": {"fp64_dmma_tflops": dmma_peak, "int8_tcgen05_tops": i8_peak},
        "kernels": kernels,
        "gpu_launches": int(launches_all),
        "clocks": clock_info,
        "e2e": e2e,
    }
    if args.workload == "cfg3":
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                "sample": "omitted: one iteration of the reference at N=1e5, D=100 takes ~20 s"}
    elif not args.no_cpu_baseline and world == 1:
        cb = run_cpu_arm(args.workload, args.sampler, 1, 1, args.cpu_baseline_iters // 2, 1)
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
