"""mMALA / simplified mMALA throughput probe (run on the GPU box): python scripts/perf_probe_mmala.py [shape] [chains]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import riemannhamiltonianmontecarlo_b200 as r
from riemannhamiltonianmontecarlo_b200.engine import ess_batched

shape = sys.argv[1] if len(sys.argv) > 1 else "german"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
xx, t = r.datasets.shaped(shape)
N, D = xx.shape
WARM, ITERS = 300, 400
for simplified in (False, True):
    data = r.LogisticData(xx, t)
    s = r.MMALASampler(data, C, 1.0, simplified)
    s.set_philox(4321, 0)
    s.set_samples(ITERS, WARM)
    s.run(WARM); torch.cuda.synchronize()
    t0 = time.time(); s.run(WARM + ITERS); torch.cuda.synchronize(); dt = time.time() - t0
    st = s.state()
    ess = ess_batched(s.samples, ITERS - 1)                  # (C, D)
    min_ess = float(torch.nan_to_num(ess, nan=0.0).sum(dim=0).min().item())
    print(f"{shape} C={C} {'simplified ' if simplified else ''}mMALA eps=1: {dt / ITERS * 1e3:.3f} ms/iteration -> "
          f"{C * ITERS / dt / 1e6:.2f} M chain-iterations/s, {min_ess / dt / 1e3:.0f} k min-ESS/s, "
          f"accept {st['accepted'].sum() / st['iters'].sum():.3f}")
    data.close()
