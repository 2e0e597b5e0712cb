#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py [--long]

Every fixture is produced by ``/root/reference/code/{rmhmc,hmc,tools}.py`` driven by a
host-supplied RNG tape (oracle/ref_live.py).  While generating, the CPU oracle
(oracle/blr_oracle.py) is asserted bit-identical to the reference on the same tape;
``tests/test_oracle_golden.py`` repeats that check against the committed files, and the
``-m gpu`` tests compare the CUDA path with the same files.

``--long`` additionally regenerates ``posterior_*.npz`` (reference run with the default
6000/1000 iterations; minutes of CPU).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import blr_oracle as bo  # noqa: E402
from oracle import ref_live  # noqa: E402
from riemannhamiltonianmontecarlo_b200 import datasets  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_STEPS = 6


def _flat(x):
    return np.asarray(x, dtype=float).reshape(-1)


def rmhmc_fixture(name, xx, t, seeds, n_iter, burn_in, n_leapfrog, step_size, n_fixed):
    d = xx.shape[1]
    n_chain = len(seeds)
    tapes = [bo.make_tape(n_iter, d, s) for s in seeds]
    samples = np.zeros((n_chain, n_iter - burn_in, d))
    theta_steps = np.full((n_chain, n_iter, n_leapfrog, d), np.nan)
    mom_end = np.zeros((n_chain, n_iter, d))
    theta_end = np.zeros((n_chain, n_iter, d))
    mom0 = np.zeros((n_chain, n_iter, d))
    n_steps = np.zeros((n_chain, n_iter), dtype=np.int64)
    direction = np.zeros((n_chain, n_iter), dtype=np.int64)
    h_cur = np.zeros((n_chain, n_iter))
    h_prop = np.zeros((n_chain, n_iter))
    ratio = np.zeros((n_chain, n_iter))
    used_u = np.zeros((n_chain, n_iter), dtype=bool)
    accepted = np.zeros((n_chain, n_iter), dtype=bool)
    for c, tape in enumerate(tapes):
        w_ref, info = ref_live.run_rmhmc(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size, n_fixed)
        w_orc, oinfo = bo.rmhmc_chain(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size, n_fixed, record=True)
        assert np.array_equal(w_ref[1:], w_orc[1:]), f"oracle != reference on chain {c}"
        samples[c, 1:] = w_ref[1:]        # row 0 is uninitialised in the reference; stored as 0
        for e in info["steps"]:
            theta_steps[c, int(e["IterationNum"]), int(e["StepNum"])] = _flat(e["wNew"])
        for it, (e_end, e_it) in enumerate(zip(info["ends"], info["iters"])):
            mom_end[c, it] = _flat(e_end["ProposedMomentum"])
            theta_end[c, it] = _flat(e_end["wNew"])
            mom0[c, it] = _flat(e_end["OriginalMomentum"])
            n_steps[c, it] = int(_flat(e_end["RandomStep"])[0])
            direction[c, it] = int(_flat(e_end["TimeStep"])[0])
            h_cur[c, it] = _flat(e_it["CurrentH"])[0]
            h_prop[c, it] = _flat(e_it["ProposedH"])[0]
            ratio[c, it] = _flat(e_it["Ratio"])[0]
            rec = oinfo["records"][it]
            assert rec.h_current == h_cur[c, it] and rec.h_proposed == h_prop[c, it]
            assert rec.n_steps == n_steps[c, it] and rec.direction == direction[c, it]
            for s in range(rec.n_steps):
                assert np.array_equal(rec.theta_steps[s], theta_steps[c, it, s])
            assert np.array_equal(rec.mom_steps[-1] if rec.n_steps else rec.momentum0[:, 0], mom_end[c, it])
        used_u[c] = info["uniform_used"]
        accepted[c] = oinfo["accepted"]
        assert np.array_equal(used_u[c], [r.used_uniform for r in oinfo["records"]])
    stacked = bo.stack_tapes(tapes)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(
        path, xx=xx, t=t, seeds=np.array(seeds), n_iter=n_iter, burn_in=burn_in, n_leapfrog=n_leapfrog,
        step_size=step_size, n_fixed=n_fixed, z=stacked["z"], u_step=stacked["u_step"],
        z_dir=stacked["z_dir"], u_acc=stacked["u_acc"], samples=samples, theta_steps=theta_steps,
        theta_end=theta_end, mom_end=mom_end, mom0=mom0, n_steps=n_steps, direction=direction,
        h_current=h_cur, h_proposed=h_prop, ratio=ratio, used_uniform=used_u, accepted=accepted)
    print(f"{name}: {n_chain} chains x {n_iter} its, accept {accepted.mean():.2f}, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def hmc_fixture(name, xx, t, seeds, n_iter, burn_in, n_leapfrog, step_size):
    d = xx.shape[1]
    tapes = [bo.make_tape(n_iter, d, s) for s in seeds]
    samples = np.zeros((len(seeds), n_iter - burn_in, d))
    ratio = np.zeros((len(seeds), n_iter))
    theta_end = np.zeros((len(seeds), n_iter, d))
    mom_end = np.zeros((len(seeds), n_iter, d))
    accepted = np.zeros((len(seeds), n_iter), dtype=bool)
    for c, tape in enumerate(tapes):
        w_ref, info = ref_live.run_hmc(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size)
        w_orc, oinfo = bo.hmc_chain(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size, record=True)
        assert np.array_equal(w_ref, w_orc), f"HMC oracle != reference on chain {c}"
        samples[c] = w_ref
        for it, e in enumerate(info["iters"]):
            ratio[c, it] = _flat(e["Ratio"])[0]
            theta_end[c, it] = _flat(e["wNew"])
            mom_end[c, it] = _flat(e["ProposedMomentum"])
            assert oinfo["records"][it]["ratio"] == ratio[c, it]
        accepted[c] = oinfo["accepted"]
    stacked = bo.stack_tapes(tapes)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, xx=xx, t=t, seeds=np.array(seeds), n_iter=n_iter, burn_in=burn_in,
                        n_leapfrog=n_leapfrog, step_size=step_size, z=stacked["z"],
                        u_step=stacked["u_step"], u_acc=stacked["u_acc"], samples=samples,
                        ratio=ratio, theta_end=theta_end, mom_end=mom_end, accepted=accepted)
    print(f"{name}: accept {accepted.mean():.2f}, {os.path.getsize(path) / 1024:.0f} KiB")


def iwls_fixture(name, xx, t, seeds, n_iter, burn_in):
    d = xx.shape[1]
    tapes = [bo.make_tape(n_iter, d, s) for s in seeds]
    samples = np.zeros((len(seeds), n_iter - burn_in, d))
    ratio = np.zeros((len(seeds), n_iter))
    proposals = np.zeros((len(seeds), n_iter, d))
    accepted = np.zeros((len(seeds), n_iter), dtype=bool)
    used_u = np.zeros((len(seeds), n_iter), dtype=bool)
    for c, tape in enumerate(tapes):
        w_ref, info = ref_live.run_iwls(xx, t, tape, n_iter, burn_in)
        w_orc, oinfo = bo.iwls_chain(xx, t, tape, n_iter, burn_in, record=True)
        assert np.array_equal(w_ref, w_orc), f"IWLS oracle != reference on chain {c}"
        samples[c] = w_ref
        for it, e in enumerate(info["iters"]):
            ratio[c, it] = _flat(e["ratio"])[0]
            proposals[c, it] = _flat(e["beta_new"])
            assert oinfo["records"][it]["ratio"] == ratio[c, it]
        accepted[c] = oinfo["accepted"]
        used_u[c] = info["uniform_used"]
        assert np.array_equal(used_u[c], [r["used_uniform"] for r in oinfo["records"]])
    stacked = bo.stack_tapes(tapes)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, xx=xx, t=t, seeds=np.array(seeds), n_iter=n_iter, burn_in=burn_in, z=stacked["z"],
                        u_acc=stacked["u_acc"], samples=samples, ratio=ratio, proposals=proposals, accepted=accepted,
                        used_uniform=used_u)
    print(f"{name}: accept {accepted.mean():.2f}, {os.path.getsize(path) / 1024:.0f} KiB")


def tools_fixture():
    tools = ref_live.load_reference()["tools"]
    rng = np.random.default_rng(2024)
    n, phis = 5000, [0.0, 0.5, 0.9, 0.99, -0.4, -0.9]
    x = np.zeros((n, len(phis)))
    e = rng.standard_normal((n, len(phis)))
    for j, phi in enumerate(phis):
        for i in range(1, n):
            x[i, j] = phi * x[i - 1, j] + e[i, j]
    ess_full = tools.CalculateESS(x, n - 1)
    ess_599 = tools.CalculateESS(x[:599], 598)
    ess_lag50 = tools.CalculateESS(x, 50)
    ac_200 = np.stack([tools.ac(x[:, j], 200) for j in range(x.shape[1])], axis=1)
    assert np.array_equal(ess_full, bo.ess(x, n - 1)) and np.array_equal(ess_599, bo.ess(x[:599], 598))
    assert np.array_equal(ess_lag50, bo.ess(x, 50))
    w = rng.standard_normal((15, 1))
    lnp = tools.LogNormPDF(np.zeros((1, 15)), w, 100)
    assert lnp == bo.log_norm_pdf(np.zeros((1, 15)), w, 100)
    np.savez_compressed(os.path.join(HERE, "tools_ess.npz"), x=x, phis=np.array(phis), ess_full=ess_full,
                        ess_599=ess_599, ess_lag50=ess_lag50, ac_200=ac_200, lnp_w=w, lnp=lnp,
                        nextpow2=np.array([[i, tools.nextpow2(i)] for i in (1, 2, 3, 599, 4096, 5000)]))
    print("tools_ess: ESS", ess_full.ravel().round(1))


def posterior_fixture(name, xx, t, seed, n_iter=6000, burn_in=1000, n_fixed=6):
    tape = bo.make_tape(n_iter, xx.shape[1], seed)
    w_ref, info = ref_live.run_rmhmc(xx, t, tape, n_iter, burn_in, 6, 0.5, n_fixed, trace=False)
    tools = ref_live.load_reference()["tools"]
    s = w_ref[1:]
    ess_ref = tools.CalculateESS(s, s.shape[0] - 1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, n_iter=n_iter, burn_in=burn_in,
                        n_fixed=n_fixed, mean=s.mean(axis=0), var=s.var(axis=0, ddof=1),
                        ess=ess_ref[:, 0], n_samples=s.shape[0], shape=np.array(xx.shape))
    print(f"{name}: mean[:3] {s.mean(axis=0)[:3]}, minESS {ess_ref.min():.0f}/{s.shape[0]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true")
    ap.add_argument("--only-iwls", action="store_true", help="regenerate only the IWLS fixtures")
    ap.add_argument("--only-hmc-german", action="store_true", help="generate only the German-shaped HMC fixture")
    args = ap.parse_args()
    assert ref_live.available(), "needs /root/reference"

    xa, ta = datasets.shaped("australian")
    xg, tg = datasets.shaped("german")
    if args.only_hmc_german:
        hmc_fixture("hmc_german_shaped", xg, tg, seeds=[901, 902], n_iter=24, burn_in=4, n_leapfrog=100, step_size=0.05)
        return
    if args.only_iwls:
        xp, tp = datasets.load_csv(os.path.join(ref_live.REFERENCE_CODE, "data", "pima.csv"))
        iwls_fixture("iwls_australian_shaped", xa, ta, seeds=[701, 702, 703], n_iter=30, burn_in=6)
        iwls_fixture("iwls_pima_real", xp, tp, seeds=[801, 802], n_iter=30, burn_in=6)
        return
    rmhmc_fixture("rmhmc_australian_shaped", xa, ta, seeds=[101, 102, 103, 104, 105, 106],
                  n_iter=24, burn_in=4, n_leapfrog=6, step_size=0.5, n_fixed=6)
    rmhmc_fixture("rmhmc_german_shaped", xg, tg, seeds=[201, 202, 203],
                  n_iter=12, burn_in=2, n_leapfrog=6, step_size=0.5, n_fixed=6)
    xr, tr = datasets.load_csv(os.path.join(ref_live.REFERENCE_CODE, "data", "german.csv"), relabel_12=True)
    rmhmc_fixture("rmhmc_german_real", xr, tr, seeds=[301, 302],
                  n_iter=10, burn_in=2, n_leapfrog=6, step_size=0.5, n_fixed=4)
    xp, tp = datasets.load_csv(os.path.join(ref_live.REFERENCE_CODE, "data", "pima.csv"))
    rmhmc_fixture("rmhmc_pima_real", xp, tp, seeds=[401, 402, 403, 404],
                  n_iter=30, burn_in=5, n_leapfrog=6, step_size=0.5, n_fixed=4)
    hmc_fixture("hmc_australian_shaped", xa, ta, seeds=[501, 502, 503], n_iter=40, burn_in=8,
                n_leapfrog=100, step_size=0.1)
    hmc_fixture("hmc_pima_real", xp, tp, seeds=[601, 602], n_iter=40, burn_in=8,
                n_leapfrog=100, step_size=0.1)
    # D = 25 = 8 k + 1: the fused HMC kernel's FMA path for the last parameter (csrc/hmc_fused.cuh); BLR_hmc.m:72 step size
    hmc_fixture("hmc_german_shaped", xg, tg, seeds=[901, 902], n_iter=24, burn_in=4, n_leapfrog=100, step_size=0.05)
    iwls_fixture("iwls_australian_shaped", xa, ta, seeds=[701, 702, 703], n_iter=30, burn_in=6)
    iwls_fixture("iwls_pima_real", xp, tp, seeds=[801, 802], n_iter=30, burn_in=6)
    tools_fixture()
    if args.long:
        posterior_fixture("posterior_australian_shaped", xa, ta, seed=7001)
        posterior_fixture("posterior_german_shaped", xg, tg, seed=7002)


if __name__ == "__main__":
    main()
