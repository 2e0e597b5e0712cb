"""B200-native batched Riemann-manifold HMC for Bayesian logistic regression.

Drop-in entry points with the reference's names (code/rmhmc.py, code/hmc.py, code/tools.py):
``RMHMC``, ``HMC``, ``LogNormPDF``, ``nextpow2``, ``ac``, ``CalculateESS``; batched engines
``RMHMCSampler`` / ``HMCSampler`` / ``rmhmc_batched`` / ``hmc_batched`` (and ``mMALA`` / ``MMALASampler``, MATLAB-only
in the reference) on top of the C ABI in
``include/rmhmc_b200.h`` (``librmhmc_b200.so``, sm_100a only, no CPU fallback).
"""
from . import datasets  # noqa: F401
from ._capi import RmhmcError  # noqa: F401
from .engine import HMCSampler, LogisticData, MMALASampler, RMHMCSampler, autocorr_batched, ess_batched, rhat_batched  # noqa: F401
from .hmc import HMC, hmc_batched  # noqa: F401
from .iwls import iwls, iwls_batched  # noqa: F401
from .mmala import mMALA, mmala_batched  # noqa: F401
from .rmhmc import RMHMC, rmhmc_batched  # noqa: F401
from .tools import CalculateESS, LogNormPDF, ac, nextpow2  # noqa: F401

__all__ = ["RMHMC", "HMC", "mMALA", "LogNormPDF", "nextpow2", "ac", "CalculateESS", "RMHMCSampler", "HMCSampler",
           "MMALASampler", "LogisticData", "rmhmc_batched", "hmc_batched", "mmala_batched", "iwls", "iwls_batched", "ess_batched", "rhat_batched", "datasets", "RmhmcError"]
