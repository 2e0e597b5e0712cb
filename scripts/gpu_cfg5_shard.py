"""BASELINE.json configs[4] at per-GPU scale: synthetic N = 1.25 M rows PER RANK, D = 64, 64 chains, X row-sharded over
the ranks, every build / pass all-reduced with NCCL inside the library.  Launch under torchrun (2 ranks = N 2.5 M):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/gpu_cfg5_shard.py
Rows are generated per shard on the host with the SURVEY.md 8d formula (seed 1237 + rank); prints ms per round and the
per-kernel shares of rank 0, and checks that all ranks hold bit-identical chain states."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riemannhamiltonianmontecarlo_b200 as r  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N_PER, D, C = int(os.environ.get("CFG5_ROWS", 1_250_000)), 64, 64
rng = np.random.default_rng(1237 + rank)
z = rng.standard_normal((N_PER, D - 1))
x = np.empty_like(z)
x[:, 0] = z[:, 0]
for j in range(1, D - 1):
    x[:, j] = 0.3 * x[:, j - 1] + np.sqrt(1 - 0.09) * z[:, j]
x = (x - x.mean(0)) / x.std(0)                       # per-shard standardisation (timing probe; parity uses tests/multi_gpu_row_shard.py)
xx = np.hstack([np.ones((N_PER, 1)), x])
beta = np.random.default_rng(99).normal(0, 0.05, (D, 1))
t = (rng.random(N_PER) < 1 / (1 + np.exp(-(xx @ beta)[:, 0]))).astype(np.float64)
data = r.LogisticData(xx, t, device=f"cuda:{local}", row_shard=(rank, world) if world > 1 else None,
                      metric=os.environ.get("CFG5_METRIC") or None)
s = r.RMHMCSampler(data, C, 6, 0.02, 6)
s.set_philox(5, 0)
s.advance(1); torch.cuda.synchronize()
s.profile(True)
R = 3
t0 = time.time(); s.advance(R); torch.cuda.synchronize(); dt = time.time() - t0
prof = s.profile_read()
st = s.state()
same = True
if world > 1:
    th = torch.from_numpy(st["theta"]).cuda()
    ref = th.clone(); dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, th))
    flag = torch.tensor([1.0 if same else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN); same = bool(flag.item() == 1.0)
if rank == 0:
    P2 = D * (D + 1) // 2
    P2p, F = (P2 + 7) // 8 * 8, 6
    ar_ms, ar_n = prof.get("allreduce", (0.0, 0))
    out = {"metric_build": data.metric_mode, "rows_total": N_PER * world, "rows_per_rank": N_PER, "dim": D, "chains": C, "ranks": world, "partials": data.partials_mode,
           "ms_per_round": dt / R * 1e3, "ranks_bit_identical": same,
           "allreduce": {"launches_per_round": ar_n / R, "ms_per_round": ar_ms / R, "share_of_round": ar_ms / (dt * 1e3),
                         "message_bytes": {"metric_iterate (G)": C * P2p * 8, "metric_closing (G | X^T(t-p) | loglik)": C * (P2p + D + 1) * 8,
                                           "quad pass": C * D * 8, "pair pass (quad | trace)": 2 * C * D * 8},
                         "per_leapfrog_step": f"{F - 1} x G + 1 x closing + {F} x quad + 1 x pair"},
           "kernels": {k: {"ms_avg": ms / max(n, 1), "launches": n, "ms_per_round": ms / R} for k, (ms, n) in prof.items() if n},
           "metric_tflops_per_rank": 2.0 * C * N_PER * P2 / (prof["metric_fp"][0] / max(prof["metric_fp"][1], 1) * 1e-3) / 1e12}
    print(json.dumps(out))
data.close()
if world > 1:
    dist.destroy_process_group()
