"""Drop-in for the reference's ``code/rmhmc.py``: ``RMHMC(XX, t, ...) -> (wSaved, TimeTaken)``.

Same signature, defaults, return shapes and stdout as rmhmc.py:13-201; the generalized leapfrog
runs in the CUDA kernels of this package (no CPU fallback).  Randomness comes from the
process-global ``np.random`` in exactly the reference's consumption order
(``randn(1,D)`` -> ``rand()`` -> ``randn()`` -> [``rand()`` iff ``Ratio <= 0``], rmhmc.py:80,89,90,181),
so ``np.random.seed(s); RMHMC(...)`` follows the same chain as the reference (to floating-point
round-off: ~1e-12 per step, see tests/test_parity_gpu.py).

``rmhmc_batched`` is the many-chain entry point the benchmark uses (Philox draws on the device).
"""
from __future__ import annotations

import timeit

import numpy as np

from .engine import LogisticData, RMHMCSampler

ALPHA = 100  # rmhmc.py:19


def RMHMC(XX, t, NumOfIterations=6000, BurnIn=1000, NumOfLeapFrogSteps=6, StepSize=0.5, NumOfNewtonSteps=4,
          *, device="cuda:0", verbose=True):
    """RIEMANNIAN HAMILTONIAN MONTE CARLO -- one chain, reference semantics (rmhmc.py:13)."""
    XX = np.asarray(XX, dtype=np.float64)
    N, D = XX.shape
    n_saved = NumOfIterations - BurnIn
    wSaved = np.zeros((n_saved, D))           # reference: np.empty; row 0 is never written (rmhmc.py:28,190)
    data = LogisticData(XX, t, alpha=ALPHA, device=device)
    sampler = RMHMCSampler(data, 1, NumOfLeapFrogSteps, StepSize, NumOfNewtonSteps)
    sampler.set_samples(max(n_saved, 1), BurnIn)
    torch = data.torch
    # one-iteration tape window, refilled from np.random before every iteration
    z_d = torch.empty(1, 1, D, dtype=torch.float64, device=data.device)
    us_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    zd_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    ua_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    sampler._keep["tape"] = [z_d, us_d, zd_d, ua_d]
    host = torch.empty(D + 3, dtype=torch.float64).pin_memory()
    stage = torch.empty(D + 3, dtype=torch.float64, device=data.device)
    sampler.set_trace(1)
    flags = sampler.trace["flags"]
    Proposed = Accepted = 0
    start = None
    from . import _capi
    from ctypes import c_void_p
    lib, h = sampler._lib, sampler.h
    for IterationNum in range(NumOfIterations):
        if (IterationNum + 1) % 50 == 0 and verbose:                       # rmhmc.py:39-45
            print('{} iterations completed.'.format(IterationNum + 1))
            print('Acceptance: {}'.format(Accepted / Proposed))
            Accepted = 0
            Proposed = 0
        Proposed += 1
        z = np.random.randn(1, D)                                          # rmhmc.py:80
        u_step = np.random.rand()                                          # rmhmc.py:89
        z_dir = np.random.randn()                                          # rmhmc.py:90
        rng_state = np.random.get_state()
        u_acc = np.random.rand()                                           # rmhmc.py:181 (speculative)
        host[:D] = torch.from_numpy(z[0])
        host[D], host[D + 1], host[D + 2] = u_step, z_dir, u_acc
        stage.copy_(host, non_blocking=True)
        z_d.view(-1).copy_(stage[:D]); us_d.view(-1).copy_(stage[D:D + 1])
        zd_d.view(-1).copy_(stage[D + 1:D + 2]); ua_d.view(-1).copy_(stage[D + 2:D + 3])
        _capi.check(lib.rmhmc_set_tape(h, IterationNum, 1, c_void_p(z_d.data_ptr()), c_void_p(us_d.data_ptr()),
                                       c_void_p(zd_d.data_ptr()), c_void_p(ua_d.data_ptr())), h, "set_tape")
        # trace window follows the iteration so that flags[0, 0] is this iteration's outcome
        sampler.trace_base = IterationNum
        _set_trace_base(sampler, IterationNum)
        sampler.run(IterationNum + 1)
        fl = int(flags[0, 0].item())
        if fl & 1:
            Accepted += 1
        if not (fl & 2):
            np.random.set_state(rng_state)                                 # the uniform was not consumed
        if IterationNum == BurnIn:                                         # rmhmc.py:194-196
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


def _set_trace_base(sampler, it_base):
    """Point the 1-iteration trace window at iteration ``it_base`` (drop-in single-chain driver)."""
    # the trace buffers are indexed by absolute iteration; a window of one iteration is emulated by
    # offsetting the base pointers so that index `it_base` lands on element 0
    from ctypes import c_void_p
    tr, d, L = sampler.trace, sampler.dim, sampler.n_leapfrog
    n = it_base + 1

    def off(tensor, per_iter):
        return c_void_p(tensor.data_ptr() - it_base * per_iter * tensor.element_size())

    from . import _capi
    _capi.check(sampler._lib.rmhmc_set_trace(
        sampler.h, n, off(tr["theta_steps"], L * d), off(tr["mom_end"], d), off(tr["theta_end"], d),
        off(tr["mom0"], d), off(tr["h_current"], 1), off(tr["h_proposed"], 1), off(tr["flags"], 1)),
        sampler.h, "set_trace")


def rmhmc_batched(XX, t, n_chains, NumOfIterations=6000, BurnIn=1000, NumOfLeapFrogSteps=6, StepSize=0.5,
                  NumOfNewtonSteps=4, *, seed=0, chain_offset=0, device="cuda:0", draws=None, return_device=False,
                  partials=None, metric=None, regime=None, student_t=False):
    """``n_chains`` independent RMHMC chains -> ``(samples (C, n-b, D), seconds, info)``.

    ``draws`` = dict(z, u_step, z_dir, u_acc) in the (W, C, ...) layout replays a host tape
    (parity mode); otherwise Philox(seed, chain_offset + chain) draws are generated on the device.
    Row 0 of every chain's samples is never written (zeros), as in the reference.
    ``partials``: ``"tensor"`` / ``"matrix_free"`` (see ``LogisticData.set_partials_mode``); None = library default.
    ``metric``: ``"dmma"`` / ``"i8"`` (``LogisticData.set_metric_mode``); ``regime``: kernel-variant pin (tests).
    ``student_t=True``: the Student-t kinetic energy of the MATLAB original ``BLR_RMHMC_StudentT.m`` (``draws`` then also
    needs ``z_chi`` (W, C); samples are stored from iteration ``BurnIn`` on, as in that script).
    """
    data = LogisticData(XX, t, alpha=ALPHA, device=device, partials=partials, metric=metric, regime=regime)
    sampler = RMHMCSampler(data, n_chains, NumOfLeapFrogSteps, StepSize, NumOfNewtonSteps, student_t=student_t)
    if draws is not None:
        sampler.set_tape(draws["z"], draws["u_step"], draws["z_dir"], draws["u_acc"], z_chi=draws.get("z_chi"))
    else:
        sampler.set_philox(seed, chain_offset)
    sampler.set_samples(NumOfIterations - BurnIn, BurnIn)
    torch = data.torch
    sampler.run(min(BurnIn + (0 if student_t else 1), NumOfIterations))
    torch.cuda.synchronize(data.device)
    start = timeit.default_timer()
    rounds = sampler.run(NumOfIterations)
    torch.cuda.synchronize(data.device)
    seconds = timeit.default_timer() - start
    info = sampler.state()
    info["rounds_after_burn_in"] = rounds
    out = sampler.samples if return_device else sampler.samples.cpu().numpy()
    if not return_device:
        data.close()
    else:
        info["_data"] = data          # keeps the handle (and the sample buffer) alive
    return out, seconds, info
