"""Design matrices for the logistic-regression target.

Two sources, both returning ``(XX, t)`` exactly as the reference samplers expect them
(``XX`` (N, D) float64 C-contiguous with a leading column of ones, ``t`` (N, 1) float64
in {0, 1}):

* :func:`synthetic_logistic` -- the deterministic "<dataset>-shaped" generator the
  benchmark configs are quoted on;
* :func:`load_csv` -- the reference's own preprocessing of ``code/data/*.csv``
  (main.py:23-41: label = last column, optional 1/2 -> 0/1 relabel, standardise with
  the population std, prepend the intercept).
"""
from __future__ import annotations

import numpy as np

# name -> (N, D incl. intercept, seed); BASELINE.json configs 1/4, 2, 3, 5
SHAPES = {
    "german": (1000, 25, 1234),
    "australian": (690, 15, 1235),
    "wide100": (100_000, 100, 1236),
    "tall64": (10_000_000, 64, 1237),
}


def add_intercept_standardised(x: np.ndarray) -> np.ndarray:
    """Column-standardise with ddof=0 and prepend ones (main.py:34-41)."""
    x = (x - x.mean(axis=0)) / x.std(axis=0)
    return np.ascontiguousarray(np.hstack((np.ones((x.shape[0], 1)), x)))


def synthetic_logistic(n: int, d: int, seed: int):
    """AR(1)-correlated Gaussian covariates, Bernoulli labels from a random beta.

    ``Z ~ N(0,1)``; ``X_j = 0.3 X_{j-1} + sqrt(1-0.09) Z_j``; standardise (ddof=0);
    ``XX = [1 | X]``; ``beta ~ N(0, 0.5^2)``; ``t = (U < sigmoid(XX beta))``.
    """
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, d - 1))
    x = np.empty_like(z)
    x[:, 0] = z[:, 0]
    for j in range(1, d - 1):
        x[:, j] = 0.3 * x[:, j - 1] + np.sqrt(1 - 0.09) * z[:, j]
    xx = add_intercept_standardised(x)
    beta = rng.normal(0.0, 0.5, (d, 1))
    prob = 1.0 / (1.0 + np.exp(-xx.dot(beta)))
    t = (rng.random((n, 1)) < prob).astype(np.float64)
    return xx, t


def shaped(name: str):
    """One of the named benchmark shapes, e.g. ``shaped('german')`` -> (1000x25, seed 1234)."""
    n, d, seed = SHAPES[name]
    return synthetic_logistic(n, d, seed)


# The five data sets of the reference (code/data/*.csv) as the MATLAB drivers prepare them
# (authors_code/Bayes_Log_Reg/MCMC/BLR_RMHMC.m:9-178): name -> (relabel {1,2} -> {0,1}, polynomial order of the basis).
# main.py itself only wires 'heart' and 'australian' (main.py:20-32); Ripley uses the cubic basis [1, X, X^2, X^3]
# (BLR_RMHMC.m:152-175, D = 7).
DATASETS = {"australian": (False, 1), "german": (True, 1), "heart": (True, 1), "pima": (False, 1), "ripley": (False, 3)}


def polynomial_basis(x: np.ndarray, order: int) -> np.ndarray:
    """``XX = [1, X, X.^2, ..., X.^order]`` of the column-standardised covariates (ddof = 0 as in main.py:34-37; the
    MATLAB original standardises with ddof = 1, BLR_RMHMC.m:23) -- BLR_RMHMC.m:26-31, :167-172."""
    x = (x - x.mean(axis=0)) / x.std(axis=0)
    cols = [np.ones((x.shape[0], 1))] + [x ** i for i in range(1, order + 1)]
    return np.ascontiguousarray(np.hstack(cols))


def load_dataset(name: str, data_dir: str):
    """``(XX, t)`` of one of the reference's data sets by name (case-insensitive), read from ``data_dir/<name>.csv``."""
    import os
    key = name.lower()
    if key not in DATASETS:
        raise ValueError(f"Error! Dataset must be one of {sorted(DATASETS)}.")          # main.py:31-32
    relabel, order = DATASETS[key]
    raw = np.loadtxt(os.path.join(data_dir, key + ".csv"), delimiter=",")
    t = raw[:, -1].reshape(-1, 1).copy()
    if relabel:
        t = t - 1.0
    return polynomial_basis(raw[:, :-1], order), np.ascontiguousarray(t, dtype=np.float64)


def load_csv(path: str, relabel_12: bool = False):
    """Reference preprocessing of a ``code/data/*.csv`` file (main.py:23-41).

    ``relabel_12`` maps labels {1, 2} -> {0, 1} (heart: main.py:26-27; german:
    authors_code/Bayes_Log_Reg/MCMC/BLR_RMHMC.m:52-55).
    """
    raw = np.loadtxt(path, delimiter=",")
    t = raw[:, -1].reshape(-1, 1).copy()
    if relabel_12:
        t = t - 1.0
    return add_intercept_standardised(raw[:, :-1]), np.ascontiguousarray(t, dtype=np.float64)
