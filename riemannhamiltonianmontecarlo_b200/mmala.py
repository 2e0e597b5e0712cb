"""Manifold MALA / simplified manifold MALA for the Bayesian logistic regression (Girolami & Calderhead).

The reference has these samplers only as MATLAB originals
(``code/authors_code/Bayes_Log_Reg/MCMC/BLR_mMALA.m``, ``BLR_mMALA_Simp.m``; driver ``Run_mMALA_Experiments.m``); this
module gives them the call convention of the reference's Python samplers
(``Sampler(XX, t, ...) -> (wSaved, TimeTaken)``, main.py:49-53) on top of the CUDA engine -- there is no CPU
fallback.  Defaults are the MATLAB ones (10000 iterations, 5000 burn-in, StepSize 1, BLR_mMALA.m:40-42).

``mMALA`` is the single-chain drop-in: draws come from the process-global ``np.random`` in the MATLAB loop's order
(``randn(1,D)`` -> [``rand()`` iff ``Ratio <= 0``], BLR_mMALA.m:233,285); all ``NumOfIterations - BurnIn`` rows of
``wSaved`` are written (the MATLAB loop saves every iteration after burn-in).  ``mmala_batched`` runs many chains
with Philox draws on the device (or a host tape, parity mode).
"""
from __future__ import annotations

import timeit
from ctypes import c_void_p

import numpy as np

from . import _capi
from .engine import LogisticData, MMALASampler

ALPHA = 100  # BLR_mMALA.m:16


def mmala_batched(XX, t, n_chains, NumOfIterations=10000, BurnIn=5000, StepSize=1.0, Simplified=False, *, seed=0,
                  chain_offset=0, device="cuda:0", draws=None, trace=False, metric=None):
    """``n_chains`` independent chains -> ``(samples (C, n-b, D), seconds, info)``; ``draws`` = dict(z (W,C,D),
    u_acc (W,C)) replays a host tape.  ``trace=True`` adds the per-iteration proposals / ratios / flags to ``info``."""
    data = LogisticData(XX, t, alpha=ALPHA, device=device, metric=metric)
    sampler = MMALASampler(data, n_chains, StepSize, Simplified)
    if draws is not None:
        sampler.set_tape(draws["z"], draws["u_acc"])
    else:
        sampler.set_philox(seed, chain_offset)
    sampler.set_samples(max(NumOfIterations - BurnIn, 1), BurnIn)
    if trace:
        sampler.set_trace(NumOfIterations)
    torch = data.torch
    sampler.run(min(BurnIn, NumOfIterations))
    torch.cuda.synchronize(data.device)
    start = timeit.default_timer()
    rounds = sampler.run(NumOfIterations)
    torch.cuda.synchronize(data.device)
    seconds = timeit.default_timer() - start
    info = sampler.state()
    info["rounds_after_burn_in"] = rounds
    if trace:
        tr = sampler.trace_numpy()
        info["proposals"], info["ratio"] = tr["theta_end"], tr["h_proposed"]
        info["accepted_flags"], info["used_uniform"] = tr["accepted"], tr["used_uniform"]
    out = sampler.samples.cpu().numpy()
    data.close()
    return out, seconds, info


def mMALA(XX, t, NumOfIterations=10000, BurnIn=5000, StepSize=1.0, Simplified=False, *, device="cuda:0", verbose=True):
    """One chain driven by the global ``np.random`` (see the module docstring)."""
    XX = np.asarray(XX, dtype=np.float64)
    N, D = XX.shape
    n_saved = NumOfIterations - BurnIn
    wSaved = np.zeros((n_saved, D))
    data = LogisticData(XX, t, alpha=ALPHA, device=device)
    sampler = MMALASampler(data, 1, StepSize, Simplified)
    sampler.set_samples(max(n_saved, 1), BurnIn)
    torch = data.torch
    z_d = torch.empty(1, 1, D, dtype=torch.float64, device=data.device)
    ua_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    sampler._keep["tape"] = [z_d, ua_d]
    host = torch.empty(D + 1, dtype=torch.float64).pin_memory()
    stage = torch.empty(D + 1, dtype=torch.float64, device=data.device)
    sampler.set_trace(1)
    tr = sampler.trace
    flags = tr["flags"]
    lib, h = sampler._lib, sampler.h
    Proposed = Accepted = 0
    start = timeit.default_timer() if BurnIn <= 0 else None
    es = 8
    for IterationNum in range(NumOfIterations):
        if (IterationNum + 1) % 1000 == 0 and IterationNum + 1 < BurnIn and verbose:     # BLR_mMALA.m:218-221
            print('{} iterations completed.'.format(IterationNum + 1))
        Proposed += 1
        z = np.random.randn(1, D)                                                        # BLR_mMALA.m:233
        rng_state = np.random.get_state()
        u_acc = np.random.rand()                                                         # BLR_mMALA.m:285 (speculative)
        host[:D] = torch.from_numpy(z[0])
        host[D] = u_acc
        stage.copy_(host, non_blocking=True)
        z_d.view(-1).copy_(stage[:D]); ua_d.view(-1).copy_(stage[D:D + 1])
        _capi.check(lib.mmala_set_tape(h, IterationNum, 1, c_void_p(z_d.data_ptr()), c_void_p(ua_d.data_ptr())), h,
                    "mmala_set_tape")
        # one-iteration trace window: offset the bases so that index IterationNum lands on element 0
        _capi.check(lib.rmhmc_set_trace(
            h, IterationNum + 1, c_void_p(tr["theta_steps"].data_ptr() - IterationNum * D * es),
            c_void_p(tr["mom_end"].data_ptr() - IterationNum * D * es),
            c_void_p(tr["theta_end"].data_ptr() - IterationNum * D * es),
            c_void_p(tr["mom0"].data_ptr() - IterationNum * D * es),
            c_void_p(tr["h_current"].data_ptr() - IterationNum * es),
            c_void_p(tr["h_proposed"].data_ptr() - IterationNum * es),
            c_void_p(flags.data_ptr() - IterationNum * 4)), h, "set_trace")
        sampler.run(IterationNum + 1)
        fl = int(flags[0, 0].item())
        if fl & 1:
            Accepted += 1
        if not (fl & 2):
            np.random.set_state(rng_state)                                               # the uniform was not consumed
        if IterationNum % 100 == 0 and IterationNum + 1 < BurnIn and verbose:            # BLR_mMALA.m:302-308
            print(Accepted / Proposed)
            Proposed = Accepted = 0
        if IterationNum + 1 == BurnIn:                                                   # BLR_mMALA.m:318-321
            if verbose:
                print('Burn-in complete, now drawing posterior samples.')
            start = timeit.default_timer()
    if start is None:
        raise UnboundLocalError("cannot access local variable 'start' where it is not associated with a value")
    torch.cuda.synchronize(data.device)
    TimeTaken = timeit.default_timer() - start
    if n_saved > 0:
        wSaved[:] = sampler.samples[0, :n_saved].cpu().numpy()
    data.close()
    return wSaved, TimeTaken
