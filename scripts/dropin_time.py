import sys, time
sys.path.insert(0, '.')
import numpy as np
import riemannhamiltonianmontecarlo_b200 as r
xx, t = r.datasets.shaped("german")
np.random.seed(1)
t0 = time.time()
w, tt = r.RMHMC(xx, t, 1200, 200, 6, 0.5, 6, verbose=False)
print("drop-in RMHMC german 1200 iterations: wall", round(time.time() - t0, 2), "s; TimeTaken (1000 post-burn-in)", round(tt, 2), "s ->", round(1000 / tt, 1), "it/s")
ess = r.CalculateESS(w[1:], w.shape[0] - 2)
print("min ESS", float(ess.min()), "-> min-ESS/s", float(ess.min()) / tt)
np.random.seed(1)
t0 = time.time()
w, tt = r.HMC(xx, t, 600, 100, 100, 0.01, verbose=False)
print("drop-in HMC 600 iterations L<=100: TimeTaken", round(tt, 2), "s")
