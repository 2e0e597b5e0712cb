"""Drop-in for the reference's ``code/iwls.py``: ``iwls(XX, t, alpha=100, max_iter=10000, burn_in=5000) -> (beta_saved, time)``.

Iterated-weighted-least-squares Metropolis-Hastings (Gamerman 1997): the proposal N(cov X^T W z, cov) with
cov = (X^T W X + I/alpha)^-1 and z = X beta + W^-1 (t - p) (iwls.py:28-35) equals N(beta + G^-1 grad, G^-1) with the Fisher
metric G and the gradient of the log joint, so it runs on the metric / Cholesky / Metropolis kernels of the manifold-MALA
path (csrc/mmala_kernels.cuh, variant 2) -- one metric build per iteration, no CPU fallback.

The reference draws with ``np.random.multivariate_normal`` (an SVD factor).  The single-chain drop-in follows the
process-global ``np.random`` in the reference's order (``multivariate_normal`` -> [``uniform`` iff ratio <= 0]): the host
draws ``beta_new`` from the device's current (mean, cov) with that very call and hands the engine the standard normals
``z = L^-1 (beta_new - mean)`` that reproduce it through the Cholesky factor.  ``iwls_batched`` is the many-chain
entry point (Philox draws on the device).
"""
from __future__ import annotations

import timeit
from ctypes import c_void_p

import numpy as np

from . import _capi
from .engine import LogisticData, MMALASampler


def iwls_batched(XX, t, n_chains, max_iter=10000, burn_in=5000, alpha=100, *, seed=0, chain_offset=0, device="cuda:0",
                 draws=None, trace=False, metric=None):
    """``n_chains`` independent IWLS chains -> ``(samples (C, max_iter - burn_in, D), seconds, info)``; ``draws`` =
    dict(z (W,C,D), u_acc (W,C)) replays a host tape (proposal = mean + chol(cov) z)."""
    data = LogisticData(XX, t, alpha=alpha, device=device, metric=metric)
    sampler = MMALASampler(data, n_chains, 1.0, iwls=True)
    if draws is not None:
        sampler.set_tape(draws["z"], draws["u_acc"])
    else:
        sampler.set_philox(seed, chain_offset)
    sampler.set_samples(max(max_iter - burn_in, 1), burn_in)
    if trace:
        sampler.set_trace(max_iter)
    torch = data.torch
    sampler.run(min(burn_in, max_iter))
    torch.cuda.synchronize(data.device)
    start = timeit.default_timer()
    sampler.run(max_iter)
    torch.cuda.synchronize(data.device)
    seconds = timeit.default_timer() - start
    info = sampler.state()
    if trace:
        tr = sampler.trace_numpy()
        info["proposals"], info["ratio"] = tr["theta_end"], tr["h_proposed"]
        info["accepted_flags"], info["used_uniform"] = tr["accepted"], tr["used_uniform"]
    out = sampler.samples.cpu().numpy()
    data.close()
    return out, seconds, info


def iwls(XX, t, alpha=100, max_iter=10000, burn_in=5000, *, device="cuda:0", verbose=True):
    """One chain, reference semantics and stdout (iwls.py:13-89), driven by the global ``np.random``."""
    XX = np.asarray(XX, dtype=np.float64)
    n_samples, dim = XX.shape
    if verbose:
        print("--- Initialization...")
    beta_saved = np.zeros((max_iter - burn_in, dim))
    data = LogisticData(XX, t, alpha=alpha, device=device)
    sampler = MMALASampler(data, 1, 1.0, iwls=True)
    sampler.set_samples(max(max_iter - burn_in, 1), burn_in)
    torch = data.torch
    z_d = torch.empty(1, 1, dim, dtype=torch.float64, device=data.device)
    ua_d = torch.empty(1, 1, dtype=torch.float64, device=data.device)
    sampler._keep["tape"] = [z_d, ua_d]
    sampler.set_trace(1)
    tr = sampler.trace
    flags = tr["flags"]
    lib, h = sampler._lib, sampler.h
    accepted = 0
    start = None
    es = 8
    if verbose:
        print("--- Iterating...")
    for i in range(max_iter):
        if i % 1000 == 0 and verbose:
            print("Iteration %d" % i)
        if i == burn_in:
            if verbose:
                print("Burn-in complete, now drawing posterior samples.")
            start = timeit.default_timer()
        mean, chol = sampler.proposal_moments()                            # current mean and lower factor of cov, (D,), (D, D)
        beta_new = np.random.multivariate_normal(mean, chol @ chol.T)     # iwls.py:45
        rng_state = np.random.get_state()
        u_acc = np.random.uniform()                                        # iwls.py:76 (speculative)
        z = np.linalg.solve(chol, beta_new - mean)
        z_d.copy_(torch.from_numpy(z.reshape(1, 1, dim)))
        ua_d.fill_(float(u_acc))
        _capi.check(lib.mmala_set_tape(h, i, 1, c_void_p(z_d.data_ptr()), c_void_p(ua_d.data_ptr())), h, "mmala_set_tape")
        _capi.check(lib.rmhmc_set_trace(
            h, i + 1, c_void_p(tr["theta_steps"].data_ptr() - i * dim * es), c_void_p(tr["mom_end"].data_ptr() - i * dim * es),
            c_void_p(tr["theta_end"].data_ptr() - i * dim * es), c_void_p(tr["mom0"].data_ptr() - i * dim * es),
            c_void_p(tr["h_current"].data_ptr() - i * es), c_void_p(tr["h_proposed"].data_ptr() - i * es),
            c_void_p(flags.data_ptr() - i * 4)), h, "set_trace")
        sampler.run(i + 1)
        fl = int(flags[0, 0].item())
        if fl & 1:
            accepted += 1
        if not (fl & 2):
            np.random.set_state(rng_state)                                 # the uniform was not consumed
    if start is None:
        raise UnboundLocalError("cannot access local variable 'start' where it is not associated with a value")
    if verbose:
        print("--- Iterating: done.")
        print("Number of accepted samples: ", accepted)
    torch.cuda.synchronize(data.device)
    time = timeit.default_timer() - start
    if max_iter > burn_in:
        beta_saved[:] = sampler.samples[0, :max_iter - burn_in].cpu().numpy()
    data.close()
    return beta_saved, time
