"""Headless version of the reference's experiment script ``code/main.py``.

``main.py`` loads a CSV (label = last column), optionally relabels {1,2} -> {0,1}, standardises with
the population std, prepends the intercept (main.py:20-41), runs one sampler 10 times sequentially
(main.py:43-53), averages the 10 sample arrays elementwise (main.py:54-55), then -- after a
``pdb.set_trace()`` and matplotlib plots -- prints the ESS summary of the AVERAGED chain and
"Time per Min ESS" (main.py:70-79).  Here the repeats are the chains of ONE batched GPU run, the
debugger/plots are dropped, and the same summary is returned and printed.

    python -m riemannhamiltonianmontecarlo_b200.harness data.csv --sampler rmhmc [--relabel-12]
    python -m riemannhamiltonianmontecarlo_b200.harness german --data-dir /path/to/code/data --sampler hmc

A data-set NAME (australian, german, heart, pima, ripley) instead of a CSV path applies the per-data-set preparation of
the MATLAB drivers (relabelling, Ripley's cubic basis: datasets.load_dataset) and, for HMC, their per-data-set step size
(BLR_hmc.m:36,72,108,138,168) unless --step-size is given.
"""
from __future__ import annotations

import argparse

import numpy as np

from . import datasets
from .hmc import hmc_batched
from .iwls import iwls_batched
from .mmala import mmala_batched
from .rmhmc import rmhmc_batched
from .tools import CalculateESS

# per-dataset HMC step sizes of the MATLAB originals (authors_code/Bayes_Log_Reg/MCMC/BLR_hmc.m:36,72,108,138,168);
# the Python default 0.14 gives 0 % acceptance on the Australian and German data
HMC_STEP_SIZES = {"australian": 0.1, "german": 0.05, "heart": 0.14, "pima": 0.1, "ripley": 0.14}


def run_experiments(XX, t, sampler="rmhmc", n_experiments=10, NumOfIterations=6000, BurnIn=1000, seed=0,
                    device="cuda:0", verbose=True, **sampler_kwargs):
    """main.py:43-79 for one data set: returns a dict with the per-experiment samples and the summary."""
    if sampler == "rmhmc":
        samples, seconds, info = rmhmc_batched(XX, t, n_experiments, NumOfIterations, BurnIn, seed=seed, device=device,
                                               **sampler_kwargs)
    elif sampler == "hmc":
        samples, seconds, info = hmc_batched(XX, t, n_experiments, NumOfIterations, BurnIn, seed=seed, device=device,
                                             **sampler_kwargs)
    elif sampler in ("mmala", "mmala_simp"):                 # the MATLAB originals' samplers (BLR_mMALA.m, BLR_mMALA_Simp.m)
        samples, seconds, info = mmala_batched(XX, t, n_experiments, NumOfIterations, BurnIn, Simplified=sampler == "mmala_simp",
                                               seed=seed, device=device, **sampler_kwargs)
    elif sampler == "iwls":                                  # code/iwls.py (main.py:51)
        sampler_kwargs.pop("StepSize", None)
        samples, seconds, info = iwls_batched(XX, t, n_experiments, NumOfIterations, BurnIn, seed=seed, device=device,
                                              **sampler_kwargs)
    else:
        raise ValueError("sampler must be 'rmhmc', 'hmc', 'mmala', 'mmala_simp' or 'iwls'")
    first = 0 if sampler.startswith("mmala") or sampler == "iwls" else 1     # RMHMC / HMC never write row 0 (rmhmc.py:190, hmc.py:83)
    results_beta = samples                                   # (n_experiments, NumOfIterations-BurnIn, D), main.py:46
    avg_beta_posterior = results_beta.mean(axis=0)           # main.py:54
    ess = CalculateESS(avg_beta_posterior, avg_beta_posterior.shape[0] - 1)      # main.py:71
    per_chain = np.stack([CalculateESS(results_beta[i, first:], results_beta.shape[1] - 1 - first)[:, 0]
                          for i in range(n_experiments)])
    out = {
        "results_beta": results_beta, "avg_beta_posterior": avg_beta_posterior, "avg_time_taken": seconds,
        "ESS": ess, "Min": float(np.min(ess)), "Median": float(np.median(ess)), "Mean": float(np.mean(ess)),
        "Max": float(np.max(ess)), "TimePerMinESS": round(seconds / float(np.min(ess)), 6),
        "per_chain_ess": per_chain, "accept_rate": float(info["accepted"].sum() / max(info["iters"].sum(), 1)),
    }
    if verbose:                                              # main.py:70-79
        print('ESS')
        print('Min', out["Min"])
        print('Median', out["Median"])
        print('Mean', out["Mean"])
        print('Max', out["Max"])
        print('Time', seconds)
        print('Time per Min ESS:', out["TimePerMinESS"])
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("csv", help="path of a CSV file (label = last column) or the name of one of the reference's data sets")
    ap.add_argument("--data-dir", default="data", help="directory holding <name>.csv when a data-set name is given (main.py: data/)")
    ap.add_argument("--relabel-12", action="store_true", help="labels {1,2} -> {0,1} (heart, german)")
    ap.add_argument("--sampler", default="rmhmc", choices=["rmhmc", "hmc", "mmala", "mmala_simp", "iwls"])
    ap.add_argument("--experiments", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=6000)
    ap.add_argument("--burn-in", type=int, default=1000)
    ap.add_argument("--step-size", type=float, default=None)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    kw = {}
    if args.csv.lower() in datasets.DATASETS:
        XX, t = datasets.load_dataset(args.csv, args.data_dir)
        if args.sampler == "hmc":
            kw["StepSize"] = HMC_STEP_SIZES[args.csv.lower()]
    else:
        XX, t = datasets.load_csv(args.csv, relabel_12=args.relabel_12)
    if args.step_size is not None:
        kw["StepSize"] = args.step_size
    return run_experiments(XX, t, args.sampler, args.experiments, args.iterations, args.burn_in, args.seed, **kw)


if __name__ == "__main__":
    main()
