"""Drop-in for the reference's ``code/hmc.py``: ``HMC(XX, t, ...) -> (wSaved, TimeTaken)``.

Same signature, defaults (incl. StepSize=0.14, which gives 0 % acceptance on the Australian and
German data -- hmc.py:12), return shapes and stdout as hmc.py:12-99.  Draws come
from the global ``np.random`` in the reference's order (``randn(1,D)`` -> ``rand()`` ->
[``rand()`` iff ``Ratio <= 0``], hmc.py:41,48,77).
"""
from __future__ import annotations

import timeit
from ctypes import c_void_p

import numpy as np

from . import _capi
from .engine import HMCSampler, LogisticData

ALPHA = 100  # hmc.py:18


def HMC(XX, t, NumOfIterations=6000, BurnIn=1000, NumOfLeapFrogSteps=100, StepSize=0.14,
        *, device="cuda:0", verbose=True):
    """HAMILTONIAN MONTE CARLO -- one chain, reference semantics (hmc.py:12)."""
    XX = np.asarray(XX, dtype=np.float64)
    N, D = XX.shape
    n_saved = NumOfIterations - BurnIn
    wSaved = np.zeros((n_saved, D))                                        # hmc.py:28
    data = LogisticData(XX, t, alpha=ALPHA, device=device)
    sampler = HMCSampler(data, 1, NumOfLeapFrogSteps, StepSize)
    sampler.set_samples(max(n_saved, 1), BurnIn)
    torch = data.torch
    z_d = torch.empty(1, 1, D, dtype=torch.float64, device=data.device)
    us_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    ua_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    sampler._keep["tape"] = [z_d, us_d, ua_d]
    host = torch.empty(D + 2, dtype=torch.float64).pin_memory()
    stage = torch.empty(D + 2, dtype=torch.float64, device=data.device)
    sampler.set_trace(1)
    tr = sampler.trace
    flags = tr["flags"]
    lib, h = sampler._lib, sampler.h
    Proposed = Accepted = 0
    start = None
    for IterationNum in range(NumOfIterations):
        z = np.random.randn(1, D)                                          # hmc.py:41
        Proposed += 1
        u_step = np.random.rand()                                          # hmc.py:48
        rng_state = np.random.get_state()
        u_acc = np.random.rand()                                           # hmc.py:77 (speculative)
        host[:D] = torch.from_numpy(z[0])
        host[D], host[D + 1] = u_step, u_acc
        stage.copy_(host, non_blocking=True)
        z_d.view(-1).copy_(stage[:D]); us_d.view(-1).copy_(stage[D:D + 1]); ua_d.view(-1).copy_(stage[D + 1:D + 2])
        _capi.check(lib.hmc_set_tape(h, IterationNum, 1, c_void_p(z_d.data_ptr()), c_void_p(us_d.data_ptr()),
                                     c_void_p(ua_d.data_ptr())), h, "hmc_set_tape")
        es = 8
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
            np.random.set_state(rng_state)
        if IterationNum <= BurnIn and IterationNum % 50 == 0 and verbose:  # hmc.py:85-89
            print('{} iterations completed.'.format(IterationNum))
            print('Acceptance: {}'.format(Accepted / Proposed))
            Accepted = 0
            Proposed = 0
        if IterationNum == BurnIn:                                         # hmc.py:92-94
            if verbose:
                print('Burn-in complete, now drawing posterior samples.')
            start = timeit.default_timer()
    if start is None:
        raise UnboundLocalError("cannot access local variable 'start' where it is not associated with a value")
    torch.cuda.synchronize(data.device)
    TimeTaken = timeit.default_timer() - start
    if verbose:
        print('Time drawing posterior: {}'.format(TimeTaken))
    if n_saved > 0:
        wSaved[:] = sampler.samples[0, :n_saved].cpu().numpy()
    data.close()
    return wSaved, TimeTaken


def hmc_batched(XX, t, n_chains, NumOfIterations=6000, BurnIn=1000, NumOfLeapFrogSteps=100, StepSize=0.14,
                *, seed=0, chain_offset=0, device="cuda:0", draws=None):
    """``n_chains`` independent HMC chains -> ``(samples (C, n-b, D), seconds, info)``."""
    data = LogisticData(XX, t, alpha=ALPHA, device=device)
    sampler = HMCSampler(data, n_chains, NumOfLeapFrogSteps, StepSize)
    if draws is not None:
        sampler.set_tape(draws["z"], draws["u_step"], draws["u_acc"])
    else:
        sampler.set_philox(seed, chain_offset)
    sampler.set_samples(NumOfIterations - BurnIn, BurnIn)
    torch = data.torch
    sampler.run(min(BurnIn + 1, NumOfIterations))
    torch.cuda.synchronize(data.device)
    start = timeit.default_timer()
    rounds = sampler.run(NumOfIterations)
    torch.cuda.synchronize(data.device)
    seconds = timeit.default_timer() - start
    info = sampler.state()
    info["rounds_after_burn_in"] = rounds
    out = sampler.samples.cpu().numpy()
    data.close()
    return out, seconds, info
