"""Drop-in for the reference's ``code/tools.py``: LogNormPDF, nextpow2, ac, CalculateESS.

``CalculateESS`` (the benchmark's metric, tools.py:32-74) and ``ac`` run on the GPU through
``blr_ess_batched`` / ``blr_autocorr``; they raise without a CUDA device -- no CPU fallback.
``LogNormPDF`` and ``nextpow2`` are scalar host helpers (the samplers evaluate the Gaussian
log-prior inside their kernels; these exist for callers of the reference API).
"""
from __future__ import annotations

import numpy as np

from .engine import autocorr_batched, ess_batched


def LogNormPDF(Values, Means, Variance):
    """Sum of iid Gaussian log densities (tools.py:10-14); Values (1,D) or (D,1), Means (D,1)."""
    Values = np.asarray(Values, dtype=np.float64)
    if Values.shape[1] > 1:
        Values = Values.T
    return float(np.sum(-0.5 * np.log(2 * np.pi * Variance) - (Values - Means) ** 2 / (2 * Variance)))


def nextpow2(i):
    """Smallest power of two >= i, as a value (tools.py:16-19)."""
    n = 1
    while n < i:
        n *= 2
    return n


def ac(Series, nLag):
    """Normalised circular autocorrelation, lags 0..nLag, period nextpow2(len)+1 (tools.py:21-30), on the GPU."""
    x = np.asarray(Series, dtype=np.float64).flatten()
    return autocorr_batched(x[None, :], int(nLag)).cpu().numpy()[0]


def CalculateESS(Samples, MaxLag):
    """Geyer initial-monotone-sequence ESS per column -> (D, 1) (tools.py:32-74), on the GPU."""
    Samples = np.ascontiguousarray(Samples, dtype=np.float64)
    ess = ess_batched(Samples[None, :, :], int(MaxLag))
    return ess.cpu().numpy().reshape(-1, 1)
