"""Quick per-kernel timing probe (run on the GPU box): python scripts/perf_probe.py [shape] [chains...]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import riemannhamiltonianmontecarlo_b200 as r

import os
shape = sys.argv[1] if len(sys.argv) > 1 else "german"
chains = [int(a) for a in sys.argv[2:]] or [4096, 16384, 65536]
R = int(os.environ.get("PROBE_ROUNDS", "20"))
WARM = int(os.environ.get("PROBE_WARM", "10"))
if shape.startswith("synthetic:"):          # synthetic:N:D:seed  (BASELINE.json configs[2]: synthetic:100000:100:1236)
    _, n_, d_, seed_ = shape.split(":")
    xx, t = r.datasets.synthetic_logistic(int(n_), int(d_), int(seed_))
else:
    xx, t = r.datasets.shaped(shape)
N, D = xx.shape
P2, P3 = D * (D + 1) // 2, D * (D + 1) * (D + 2) // 6
F = 6
MODES = os.environ.get("PROBE_PARTIALS", "matrix_free,tensor").split(",")
for C, mode in [(c, m) for c in chains for m in MODES]:
    data = r.LogisticData(xx, t, partials=mode)
    s = r.RMHMCSampler(data, C, 6, 0.5, F)
    s.set_philox(1234, 0)
    s.advance(WARM); torch.cuda.synchronize()
    s.profile(True)
    t0 = time.time(); s.advance(R); torch.cuda.synchronize(); dt = time.time() - t0
    prof = s.profile_read()
    s.profile(False)
    t0 = time.time(); s.advance(R); torch.cuda.synchronize(); dt2 = time.time() - t0
    w_alg = 2.0 * N * P3 + 2.0 * F * N * P2
    print(f"{shape} C={C} partials={mode}: {dt/R*1e3:.3f} ms/round (profiled) {dt2/R*1e3:.3f} ms/round (plain) -> "
          f"{C*R/dt2/1e6:.3f} M chain-leapfrog/s, {w_alg*C*R/dt2/1e12:.2f} TF/s algorithmic")
    for k, (ms, n) in prof.items():  # noqa
        if not n:
            continue
        per = ms / max(n, 1)
        extra = ""
        if k == "partials": extra = f" -> {2.0*C*N*P3/per/1e9:.2f} TF/s"
        if k == "leverage_gemm": extra = f" -> {2.0*C*N*P2/per/1e9:.2f} TF/s"
        if k == "quad_pass": extra = f" -> {4.0*C*N*D/per/1e9:.2f} TF/s"
        if k in ("metric_fp", "metric_closing"): extra = f" -> {2.0*C*N*P2/per/1e9:.2f} TF/s (G only)"
        print(f"   {k:15s} {ms:9.2f} ms / {n:4d} launches = {per:8.4f} ms{extra}  [{ms/(dt*1e3)*100:.1f}% of wall]")
    st = s.state()
    print("   iters", st["iters"].mean(), "accept", st["accepted"].sum() / max(st["iters"].sum(), 1))
    data.close()
