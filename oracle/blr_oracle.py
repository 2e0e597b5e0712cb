"""CPU restatement of the reference's Bayesian-logistic-regression samplers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): this module is the checker
for the CUDA path and the CPU arm timed by ``bench.py``; the product never calls it.

What it restates (reference = /root/reference/code, cited as file:line):

* ``log_norm_pdf``                       tools.py:10-14
* ``next_pow2`` / ``autocorr`` / ``ess``  tools.py:16-19 / 21-30 / 32-74
* ``fisher_metric``                      rmhmc.py:51-57 (= :116-119, :134-137)
* ``metric_partials``                    rmhmc.py:64-77 (= :142-156)
* ``likelihood_gradient``                rmhmc.py:100 (= :140, hmc.py:53,61)
* ``log_joint``                          rmhmc.py:31-34 (= :166-169, hmc.py:31-34,64-67)
* ``leapfrog_step``                      rmhmc.py:96-163 (one generalized leapfrog step)
* ``rmhmc_chain``                        rmhmc.py:37-191 (one chain, whole loop)
* ``hmc_chain``                          hmc.py:38-89
* ``rhat``                               not in the reference (classic Gelman-Rubin; own spec)
* ``mmala_chain``                        MATLAB only: authors_code/Bayes_Log_Reg/MCMC/BLR_mMALA.m:159-330,
                                         BLR_mMALA_Simp.m:170-290 -- PARITY UNPINNED (no MATLAB/Octave here)

Unlike the reference, randomness is an explicit argument (a :class:`DrawTape`) so
that the oracle, the live reference (``oracle/ref_live.py`` monkeypatches
``np.random``) and the CUDA path consume identical draws.  The arithmetic keeps the
reference's NumPy operation order (same BLAS calls on same-layout operands), which
is what makes ``samples[1:]`` bit-identical to the reference in this container --
checked by ``tests/golden/make_golden.py`` when the fixtures are generated and by
``tests/test_oracle_golden.py`` against the committed fixtures.

Reference quirks that are kept on purpose: momentum drawn as
``L^T z`` with L the *lower* Cholesky factor; direction ``+1`` only if a normal draw
exceeds 0.5; exactly ``n_fixed`` fixed-point iterations, never iterate-to-tolerance;
``exp(f)/(1+exp(f))`` in the gradient but ``1/(1+exp(-f))`` in the metric; the MH
uniform is consumed only when ``Ratio <= 0``; row 0 of the sample array is never
written (NaN here, uninitialised memory in the reference); the two renormalisation
hacks on the momentum (norm > 100) and the position (norm > 10).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

ALPHA = 100.0  # prior variance, hard-coded in rmhmc.py:19 and hmc.py:18


# --------------------------------------------------------------------------- draws
@dataclass
class DrawTape:
    """Host-supplied random draws for ONE chain, in the reference's consumption order.

    Per MCMC iteration the reference consumes (rmhmc.py:80,89,90,181):
    ``randn(1,D)`` -> ``rand()`` -> ``randn()`` -> [``rand()`` iff ``Ratio <= 0``].
    HMC (hmc.py:41,48,78) consumes ``randn(1,D)`` -> ``rand()`` -> [``rand()``].
    """

    z: np.ndarray       # (n_iter, D) standard normals for the momentum draw
    u_step: np.ndarray  # (n_iter,)   uniform that picks RandomStep
    z_dir: np.ndarray   # (n_iter,)   normal that picks the integration direction (RMHMC only)
    u_acc: np.ndarray   # (n_iter,)   Metropolis uniform (consumed only if Ratio <= 0)

    @property
    def n_iter(self) -> int:
        return self.z.shape[0]


def make_tape(n_iter: int, dim: int, seed: int) -> DrawTape:
    """Deterministic tape for one chain: ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    return DrawTape(
        z=rng.standard_normal((n_iter, dim)),
        u_step=rng.random(n_iter),
        z_dir=rng.standard_normal(n_iter),
        u_acc=rng.random(n_iter),
    )


def stack_tapes(tapes: list[DrawTape]) -> dict[str, np.ndarray]:
    """Batched layout used by the C ABI: iteration-major, chain-minor.

    z[n_iter, C, D], u_step[n_iter, C], z_dir[n_iter, C], u_acc[n_iter, C].
    """
    return {
        "z": np.ascontiguousarray(np.stack([tp.z for tp in tapes], axis=1)),
        "u_step": np.ascontiguousarray(np.stack([tp.u_step for tp in tapes], axis=1)),
        "z_dir": np.ascontiguousarray(np.stack([tp.z_dir for tp in tapes], axis=1)),
        "u_acc": np.ascontiguousarray(np.stack([tp.u_acc for tp in tapes], axis=1)),
    }


# --------------------------------------------------------------------------- tools.py
def log_norm_pdf(values: np.ndarray, means: np.ndarray, variance: float) -> float:
    """Sum of iid Gaussian log densities (tools.py:10-14)."""
    if values.shape[1] > 1:
        values = values.T
    half_log = 0.5 * np.log(2 * np.pi * variance)
    return np.sum(-half_log - ((values - means) ** 2) / (2 * variance))


def next_pow2(i: int) -> int:
    """Smallest power of two >= i, returned as a VALUE (tools.py:16-19)."""
    n = 1
    while n < i:
        n *= 2
    return n


def autocorr(series: np.ndarray, n_lag: int) -> np.ndarray:
    """Circular FFT autocorrelation with the reference's nFFT = next_pow2(len)+1 (tools.py:21-30)."""
    x = series.flatten()
    n_fft = next_pow2(len(x)) + 1
    spec = np.fft.fft(x - np.mean(x), n_fft)
    spec = spec * np.conj(spec)
    acf = np.fft.ifft(spec)[0 : n_lag + 1]
    acf = acf / acf[0]
    return np.real(acf)


def ess(samples: np.ndarray, max_lag: int) -> np.ndarray:
    """Geyer initial-monotone-sequence ESS per column, shape (D, 1) (tools.py:32-74)."""
    max_lag = int(max_lag)
    n_samples, n_par = samples.shape
    rho = np.zeros((max_lag + 1, n_par))
    for i in range(n_par):
        rho[:, i] = autocorr(samples[:, i], max_lag)
    half = (max_lag + 1) // 2
    gamma = rho[0 : 2 * half : 2] + rho[1 : 2 * half : 2]  # Gamma_j = rho_2j + rho_2j+1
    gamma = np.minimum.accumulate(gamma, axis=0)            # running minimum (:54-60)
    mono = np.zeros((n_par, 1))
    for i in range(n_par):
        n_pos = int(np.count_nonzero(gamma[:, i] > 0))
        mono[i] = -rho[0, i] + 2 * np.sum(gamma[0:n_pos, i])
        if mono[i] < 1:
            mono[i] = 1
    return n_samples / mono


def rhat(chains: np.ndarray) -> np.ndarray:
    """Classic Gelman-Rubin potential scale reduction; chains is (C, S, D) -> (D,).

    Not part of the reference; specified here as W = mean_c var_c (ddof=1),
    B/S = var_c(mean_c) (ddof=1), Rhat = sqrt(((S-1)/S * W + B/S) / W).
    """
    n_chain, n_samp, _ = chains.shape
    means = chains.mean(axis=1)
    w = chains.var(axis=1, ddof=1).mean(axis=0)
    b_over_s = means.var(axis=0, ddof=1)
    return np.sqrt(((n_samp - 1) / n_samp * w + b_over_s) / w)


# --------------------------------------------------------------------------- model pieces
def fisher_metric(xx: np.ndarray, w: np.ndarray, alpha: float = ALPHA):
    """f = Xw, p = sigma(f), v = p(1-p), G = X^T diag(v) X + I/alpha (rmhmc.py:51-57)."""
    n, d = xx.shape
    f = xx.dot(w)
    p = 1 / (1 + np.exp(-f))
    v = (p * (1 - p))[:, 0]
    xt_lam = np.empty((d, n))          # C-ordered (D,N), like the reference's XX.T * tile(v)
    np.multiply(xx.T, v, out=xt_lam)
    g = xt_lam.dot(xx) + np.eye(d) / alpha
    return f, p, v, g


def metric_derivative(xx: np.ndarray, p: np.ndarray, v: np.ndarray, d: int) -> np.ndarray:
    """dG/dw_d = X^T diag(v (1-2p) x_d) X (rmhmc.py:67-75)."""
    z = ((1 - 2 * p) * xx[:, d].reshape(-1, 1))[:, 0]
    z1 = v * z
    z2 = xx * z1[:, None]              # == the column loop Z2[:,a] = XX[:,a]*Z1
    return z2.T.dot(xx)


def metric_partials(xx: np.ndarray, p: np.ndarray, v: np.ndarray, inv_g: np.ndarray):
    """All D products G^-1 dG/dw_d and their traces (rmhmc.py:64-77, :142-156)."""
    d = xx.shape[1]
    inv_g_dg = np.empty((d, d, d))
    tr = np.empty((d, 1))
    for k in range(d):
        inv_g_dg[k] = inv_g.dot(metric_derivative(xx, p, v, k))
        tr[k] = np.trace(inv_g_dg[k])
    return inv_g_dg, tr


def metric_tensor(xx: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Dense T[d,a,b] = sum_n v_n (1-2p_n) x_nd x_na x_nb (the stack of all dG/dw_d)."""
    _, p, v, _ = fisher_metric(xx, w)
    return np.stack([metric_derivative(xx, p, v, k) for k in range(xx.shape[1])])


def likelihood_gradient(xx, t, w, alpha: float = ALPHA):
    """X^T (t - e^f/(1+e^f)) - w/alpha (rmhmc.py:99-100)."""
    f = xx.dot(w)
    return np.dot(xx.T, t - np.exp(f) / (1 + np.exp(f))) - np.eye(xx.shape[1]).dot(w) / alpha


def log_joint(xx, t, w, alpha: float = ALPHA):
    """f^T t - sum log(1+e^f) + log N(w; 0, alpha I) (rmhmc.py:31-34, :166-169)."""
    d = xx.shape[1]
    log_prior = log_norm_pdf(np.zeros((1, d)), w, alpha)
    f = np.dot(xx, w)
    log_lik = np.dot(f.T, t) - np.sum(np.log(1 + np.exp(f)))
    return log_lik + log_prior


def _scalar(x) -> float:
    return float(np.asarray(x).reshape(-1)[0])


def _quadratic_terms(mom, inv_g_dg, u):
    """LastTerm_d = 0.5 * mom^T (G^-1 dG_d) u (rmhmc.py:105-107, :159-161)."""
    d = mom.shape[0]
    last = np.empty((d, 1))
    for k in range(d):
        last[k] = 0.5 * mom.T.dot(inv_g_dg[k]).dot(u)
    return last


def leapfrog_step(xx, t, w_new, mom, g, inv_g, inv_g_dg, tr, sgn, step_size, n_fixed, alpha=ALPHA):
    """One generalized leapfrog step (rmhmc.py:96-163, R7-R14 in SURVEY.md 3.2).

    ``g, inv_g, inv_g_dg, tr`` are the metric quantities at ``w_new`` on entry and at the new
    position on exit.  Returns ``(w_new, mom, g, inv_g, inv_g_dg, tr, hacked)``; ``mom`` is a new
    array, the inputs are not modified.
    """
    # R7/R8: implicit momentum half-step, exactly n_fixed fixed-point iterations
    grad = likelihood_gradient(xx, t, w_new, alpha)
    pm = mom.copy()
    for _ in range(n_fixed):
        u = inv_g.dot(pm)
        last = _quadratic_terms(pm, inv_g_dg, u)
        pm = mom + sgn * step_size / 2 * (grad - 0.5 * tr + last)
    mom = pm

    # R9/R10: implicit position step
    u0 = np.linalg.solve(g, mom)
    pw = w_new.copy()
    for _ in range(n_fixed):
        _, p, v, g = fisher_metric(xx, pw, alpha)
        u = np.linalg.solve(g, mom)
        pw = w_new + sgn * step_size / 2 * (u0 + u)
    w_new = pw

    # R11: position clamp hack
    hacked = False
    if np.linalg.norm(w_new) > 10:
        w_new /= np.linalg.norm(w_new) * 3
        hacked = True

    # R12/R13: metric, gradient and partials at the new position
    _, p, v, g = fisher_metric(xx, w_new, alpha)
    inv_g = np.linalg.inv(g)
    grad = likelihood_gradient(xx, t, w_new, alpha)
    inv_g_dg, tr = metric_partials(xx, p, v, inv_g)

    # R14: explicit closing momentum half-step
    u = inv_g.dot(mom)
    last = _quadratic_terms(mom, inv_g_dg, u)
    mom += sgn * step_size / 2 * (grad - 0.5 * tr + last)
    return w_new, mom, g, inv_g, inv_g_dg, tr, hacked


def hamiltonian(xx, t, w, mom, alpha=ALPHA):
    """H = -logjoint + sum log diag chol(G) + p^T G^-1 p / 2 (rmhmc.py:166-176)."""
    _, _, _, g = fisher_metric(xx, w, alpha)
    logdet = np.sum(np.log(np.diag(np.linalg.cholesky(g))))
    return _scalar(-log_joint(xx, t, w, alpha) + logdet + mom.T.dot(np.linalg.inv(g)).dot(mom) / 2)


def leapfrog(xx, t, w, mom, sgn, n_steps, step_size, n_fixed, alpha=ALPHA):
    """``n_steps`` generalized leapfrog steps from (w, mom); returns (w, mom, H_start, H_end)."""
    d = xx.shape[1]
    w = np.array(w, dtype=float).reshape(d, 1)
    mom = np.array(mom, dtype=float).reshape(d, 1)
    h0 = hamiltonian(xx, t, w, mom, alpha)
    _, p, v, g = fisher_metric(xx, w, alpha)
    inv_g = np.linalg.inv(g)
    inv_g_dg, tr = metric_partials(xx, p, v, inv_g)
    for _ in range(n_steps):
        w, mom, g, inv_g, inv_g_dg, tr, _ = leapfrog_step(xx, t, w, mom, g, inv_g, inv_g_dg, tr, sgn, step_size,
                                                           n_fixed, alpha)
    return w[:, 0], mom[:, 0], h0, hamiltonian(xx, t, w, mom, alpha)


# --------------------------------------------------------------------------- RMHMC
@dataclass
class IterationRecord:
    """What one MCMC iteration did; filled only when ``record=True``."""

    n_steps: int = 0
    direction: int = 0
    momentum0: np.ndarray | None = None       # after the draw (and renorm hack)
    theta_steps: list = field(default_factory=list)   # position after each leapfrog step (post hack)
    mom_steps: list = field(default_factory=list)     # momentum after each closing half-step
    h_current: float = np.nan
    h_proposed: float = np.nan
    ratio: float = np.nan
    accepted: bool = False
    used_uniform: bool = False
    renorm_momentum: bool = False
    renorm_position: int = 0


def rmhmc_chain(xx, t, tape: DrawTape, n_iter=6000, burn_in=1000, n_leapfrog=6,
                step_size=0.5, n_fixed=4, alpha=ALPHA, record=False, w0=None):
    """One RMHMC chain under a draw tape; restates rmhmc.py:13-201 (R1-R19 in SURVEY.md 3.2).

    Returns ``(samples, info)``: ``samples`` is (n_iter-burn_in, D) with row 0 = NaN
    (never written by the reference, rmhmc.py:28,190); ``info`` has the final state,
    accept flags and, if ``record``, one :class:`IterationRecord` per iteration.
    """
    n, d = xx.shape
    w = np.ones((d, 1)) * 1e-3 if w0 is None else np.array(w0, dtype=float).reshape(d, 1)
    samples = np.full((n_iter - burn_in, d), np.nan)
    cur_ljl = log_joint(xx, t, w, alpha)
    accepted = np.zeros(n_iter, dtype=bool)
    steps_taken = np.zeros(n_iter, dtype=np.int64)
    records: list[IterationRecord] = []
    n_renorm_p = n_renorm_w = 0

    for it in range(n_iter):
        rec = IterationRecord() if record else None
        w_new = w.copy()

        # R2/R3: metric, inverse, Cholesky factor and partials at the current position
        _, p, v, g = fisher_metric(xx, w_new, alpha)
        inv_g = np.linalg.inv(g)
        chol0 = np.linalg.cholesky(g)
        inv_g0 = inv_g.copy()
        inv_g_dg, tr = metric_partials(xx, p, v, inv_g)

        # R4/R5: momentum = L^T z, clamp hack
        mom = np.dot(tape.z[it].reshape(1, d), chol0).T
        if np.linalg.norm(mom) > 100:
            mom /= np.linalg.norm(mom) * 25
            n_renorm_p += 1
            if rec:
                rec.renorm_momentum = True
        mom0 = mom.copy()

        # R6: trajectory length and direction
        n_steps = int(np.ceil(tape.u_step[it] * n_leapfrog))
        sgn = 1 if tape.z_dir[it] > 0.5 else -1
        steps_taken[it] = n_steps
        if rec:
            rec.n_steps, rec.direction, rec.momentum0 = n_steps, sgn, mom0.copy()

        for _step in range(n_steps):
            w_new, mom, g, inv_g, inv_g_dg, tr, hacked = leapfrog_step(
                xx, t, w_new, mom, g, inv_g, inv_g_dg, tr, sgn, step_size, n_fixed, alpha)
            if hacked:
                n_renorm_w += 1
                if rec:
                    rec.renorm_position += 1
            if rec:
                rec.theta_steps.append(w_new[:, 0].copy())
                rec.mom_steps.append(mom[:, 0].copy())

        # R15/R16: Hamiltonians
        prop_ljl = log_joint(xx, t, w_new, alpha)
        prop_logdet = np.sum(np.log(np.diag(np.linalg.cholesky(g))))
        h_prop = -prop_ljl + prop_logdet + mom.T.dot(inv_g).dot(mom) / 2
        cur_logdet = np.sum(np.log(np.diag(chol0)))
        h_cur = -cur_ljl + cur_logdet + mom0.T.dot(inv_g0).dot(mom0) / 2

        # R17: accept; the uniform is drawn only when Ratio > 0 is False
        ratio = -h_prop + h_cur
        take = bool(ratio > 0)
        used_u = False
        if not take:
            used_u = True
            take = bool(ratio > np.log(tape.u_acc[it]))
        if take:
            cur_ljl = prop_ljl
            w = w_new
            accepted[it] = True
        if rec:
            rec.h_current, rec.h_proposed = _scalar(h_cur), _scalar(h_prop)
            rec.ratio, rec.accepted, rec.used_uniform = _scalar(ratio), take, used_u
            records.append(rec)

        # R18: store (row 0 is never written)
        if it > burn_in:
            samples[it - burn_in, :] = w.T

    info = {
        "w": w[:, 0].copy(),
        "log_joint": _scalar(cur_ljl),
        "accepted": accepted,
        "steps": steps_taken,
        "renorm_momentum": n_renorm_p,
        "renorm_position": n_renorm_w,
        "records": records,
    }
    return samples, info


# --------------------------------------------------------------------------- HMC
def hmc_chain(xx, t, tape: DrawTape, n_iter=6000, burn_in=1000, n_leapfrog=100,
              step_size=0.14, alpha=ALPHA, record=False, w0=None):
    """One Euclidean-HMC chain under a draw tape; restates hmc.py:12-99.

    ``tape.z_dir`` is unused (HMC has no direction draw).  Row 0 of ``samples`` is
    zero, as in the reference (hmc.py:28,83).
    """
    n, d = xx.shape
    mass = np.eye(d)
    inv_mass = np.linalg.inv(mass)
    w = np.zeros((d, 1)) if w0 is None else np.array(w0, dtype=float).reshape(d, 1)      # w0: continue a chain (bench.py)
    samples = np.zeros((n_iter - burn_in, d))
    cur_ljl = log_joint(xx, t, w, alpha)
    accepted = np.zeros(n_iter, dtype=bool)
    steps_taken = np.zeros(n_iter, dtype=np.int64)
    records = []

    for it in range(n_iter):
        mom = np.dot(tape.z[it].reshape(1, d), mass).T
        mom0 = mom.copy()
        w_new = w.copy()
        n_steps = int(np.ceil(tape.u_step[it] * n_leapfrog))
        done = 0
        for _step in range(n_steps):
            mom += step_size / 2 * likelihood_gradient(xx, t, w_new, alpha)
            if np.sum(np.isnan(mom)) > 0:      # hmc.py:56-57
                break
            w_new += step_size * np.dot(inv_mass, mom)
            mom += step_size / 2 * likelihood_gradient(xx, t, w_new, alpha)
            done += 1
        steps_taken[it] = done
        prop_ljl = log_joint(xx, t, w_new, alpha)
        h_prop = -prop_ljl + mom.T.dot(inv_mass).dot(mom) / 2
        h_cur = -cur_ljl + mom0.T.dot(inv_mass).dot(mom0) / 2
        ratio = -h_prop + h_cur
        take = bool(ratio > 0)
        if not take:
            take = bool(ratio > np.log(tape.u_acc[it]))
        if take:
            cur_ljl = prop_ljl
            w = w_new
            accepted[it] = True
        if record:
            records.append({"n_steps": n_steps, "theta": w_new[:, 0].copy(), "mom": mom[:, 0].copy(),
                            "h_current": _scalar(h_cur), "h_proposed": _scalar(h_prop),
                            "ratio": _scalar(ratio), "accepted": take})
        if it > burn_in:
            samples[it - burn_in, :] = w.T

    info = {"w": w[:, 0].copy(), "log_joint": _scalar(cur_ljl), "accepted": accepted,
            "steps": steps_taken, "records": records}
    return samples, info


# --------------------------------------------------------------------------- mMALA (MATLAB original only)
def _mmala_terms(xx, t, w, alpha, simplified):
    """G, G^-1 and the drift pieces of BLR_mMALA.m:187-213 / :246-271 (BLR_mMALA_Simp.m:176-180,:214-224) at w."""
    d = xx.shape[1]
    f, p, v, g = fisher_metric(xx, w, alpha)
    inv_g = np.linalg.inv(g)
    first = inv_g.dot(xx.T.dot(t - np.exp(f) / (1 + np.exp(f))) - np.eye(d) * (1 / alpha) @ w)
    if simplified:
        return g, inv_g, first, None, None
    inv_g_dg, tr = metric_partials(xx, p, v, inv_g)
    second = np.empty((d, d))
    for k in range(d):
        second[:, k] = inv_g_dg[k].dot(inv_g[:, k])                    # InvGdG{d} * InvG(:, d)
    third = inv_g.dot(tr)
    return g, inv_g, first, second, third


def mmala_chain(xx, t, tape: DrawTape, n_iter=10000, burn_in=5000, step_size=1.0, alpha=ALPHA, simplified=False,
                record=False, w0=None):
    """One (simplified) manifold-MALA chain under a draw tape; restates the MATLAB original
    ``code/authors_code/Bayes_Log_Reg/MCMC/BLR_mMALA.m:159-330`` (``BLR_mMALA_Simp.m:170-290``).

    PARITY UNPINNED: the reference has no Python mMALA and MATLAB/Octave is not available, so this port cannot be
    run against the original.  Conventions: iterations are 0-based here (MATLAB IterationNum = it + 1); the sample
    after iteration ``it`` goes to row ``it - burn_in`` for ``it >= burn_in`` (MATLAB: IterationNum > BurnIn);
    per iteration the tape supplies ``z[it]`` (``randn(1,D)``, :233) and ``u_acc[it]`` (``rand``, :289, consumed only
    if Ratio <= 0 -- MATLAB ``||`` short-circuits); ``chol`` is MATLAB's upper factor R (R'R = A), so the proposal
    ``(randn(1,D) * chol(S))'`` is ``L z`` with the lower factor ``L = R'``.
    """
    n, d = xx.shape
    w = np.zeros((d, 1)) if w0 is None else np.array(w0, dtype=float).reshape(d, 1)   # :165 (w0: continue a chain)
    samples = np.zeros((n_iter - burn_in, d))
    cur_ljl = log_joint(xx, t, w, alpha)                                   # :169-172
    cur_g, cur_inv_g, cur_first, cur_second, cur_third = _mmala_terms(xx, t, w, alpha, simplified)
    accepted = np.zeros(n_iter, dtype=bool)
    records = []

    def drift(wv, first, second, third):
        mean = wv + (step_size / 2) * first                                # :227-229 / Simp :207
        if not simplified:
            mean = mean - step_size * np.sum(second, axis=1, keepdims=True) + (step_size / 2) * third
        return mean

    for it in range(n_iter):
        mean = drift(w, cur_first, cur_second, cur_third)
        l_prop = np.linalg.cholesky(step_size * cur_inv_g)
        w_new = mean + l_prop.dot(tape.z[it].reshape(d, 1))                # :233
        prop_ljl = log_joint(xx, t, w_new, alpha)                          # :236-239
        diff = mean - w_new
        p_new_old = -np.sum(np.log(np.diag(l_prop))) - 0.5 * _scalar(diff.T.dot(cur_g / step_size).dot(diff))   # :241
        g, inv_g, first, second, third = _mmala_terms(xx, t, w_new, alpha, simplified)
        mean_back = drift(w_new, first, second, third)                     # :275-277
        diff = mean_back - w
        p_old_new = -np.sum(np.log(np.diag(np.linalg.cholesky(step_size * inv_g)))) \
            - 0.5 * _scalar(diff.T.dot(g / step_size).dot(diff))           # :279
        ratio = _scalar(prop_ljl) + p_old_new - _scalar(cur_ljl) - p_new_old   # :282
        take = bool(ratio > 0)
        used_u = False
        if not take:
            used_u = True
            take = bool(ratio > np.log(tape.u_acc[it]))                    # :285
        if record:
            records.append({"theta": w_new[:, 0].copy(), "ratio": ratio, "accepted": take, "used_uniform": used_u,
                            "prop_ljl": _scalar(prop_ljl)})
        if take:
            cur_ljl = prop_ljl
            cur_g, cur_inv_g, cur_first, cur_second, cur_third = g, inv_g, first, second, third
            w = w_new
            accepted[it] = True
        if it >= burn_in:
            samples[it - burn_in, :] = w.T                                 # :313-315

    info = {"w": w[:, 0].copy(), "log_joint": _scalar(cur_ljl), "accepted": accepted, "records": records}
    return samples, info


# --------------------------------------------------------------------------- Student-t RMHMC (MATLAB original only)
def studentt_rmhmc_chain(xx, t, tape: DrawTape, z_chi, n_iter=6000, burn_in=1000, n_leapfrog=6, step_size=0.5, n_fixed=4,
                         alpha=ALPHA, record=False):
    """RMHMC with a multivariate Student-t (one degree of freedom) kinetic energy; restates the MATLAB original
    ``code/authors_code/Bayes_Log_Reg/MCMC/BLR_RMHMC_StudentT.m:205-414`` (cited as T:line).

    PARITY UNPINNED: MATLAB only, no Python original in the reference and no MATLAB / Octave here.  Differences from
    ``BLR_RMHMC.m`` / ``rmhmc.py``: the momentum draw ``mvtrnd(G, 1)'`` (T:265), the LastTerm scaling
    ``(1+D)/2 (.) / (1 + p' G^-1 p)`` (T:296, :370), the position update weights (T:311-326) and the kinetic term
    ``(1+D)/2 log(1 + p' G^-1 p)`` (T:386, :392); no renormalisation hacks (those are rmhmc.py's).
    Conventions of this port: 0-based iterations (MATLAB IterationNum = it + 1), sample of iteration ``it`` in row
    ``it - burn_in`` for ``it >= burn_in`` (T:403-405).  ``mvtrnd(C, 1)`` rescales C to a correlation matrix, draws
    ``randn(1, D) * chol(corr)`` and divides by ``sqrt(chi2rnd(1))``: the tape supplies ``z[it]`` for the normals and
    ``z_chi[it]`` with ``chi2 = z_chi^2``.  Direction ``randn > 0.5`` (T:272) from ``z_dir``, ``RandomSteps =
    ceil(rand * L)`` (T:274) from ``u_step``, the Metropolis uniform (T:398) from ``u_acc`` (consumed only if Ratio <= 0).
    """
    n, d = xx.shape
    nu = 1.0 + d
    w = np.ones((d, 1)) * 1e-3                                                   # T:207
    samples = np.zeros((n_iter - burn_in, d))
    cur_ljl = log_joint(xx, t, w, alpha)                                         # T:217-220
    accepted = np.zeros(n_iter, dtype=bool)
    steps_taken = np.zeros(n_iter, dtype=np.int64)
    records = []
    for it in range(n_iter):
        w_new = w.copy()
        _, p, v, g = fisher_metric(xx, w_new, alpha)                             # T:236-240
        inv_g = np.linalg.inv(g)
        orig_chol, orig_inv_g = np.linalg.cholesky(g), inv_g.copy()
        inv_g_dg, tr = metric_partials(xx, p, v, inv_g)                          # T:248-262
        sd = np.sqrt(np.diag(g))
        corr_l = np.linalg.cholesky(g / np.outer(sd, sd))                        # mvtrnd: chol of the correlation matrix
        mom = (corr_l.dot(tape.z[it].reshape(d, 1))) / np.sqrt(z_chi[it] ** 2)   # T:265: (z chol(corr))' / sqrt(chi2 / 1)
        mom0 = mom.copy()
        sgn = 1 if tape.z_dir[it] > 0.5 else -1                                  # T:272
        n_steps = int(np.ceil(tape.u_step[it] * n_leapfrog))                     # T:274
        steps_taken[it] = n_steps
        rec = {"theta_steps": [], "mom0": mom0[:, 0].copy(), "n_steps": n_steps, "direction": sgn}
        for _ in range(n_steps):
            grad = likelihood_gradient(xx, t, w_new, alpha)                      # T:285
            pm = mom.copy()
            for _f in range(n_fixed):                                            # T:289-300
                u = inv_g.dot(pm)
                den = 1.0 + _scalar(pm.T.dot(u))
                last = np.array([[(nu / 2) * _scalar(pm.T.dot(inv_g_dg[k]).dot(u)) / den] for k in range(d)])
                pm = mom + sgn * (step_size / 2) * (grad - 0.5 * tr + last)
            mom = pm
            u0 = np.linalg.solve(g, mom)                                         # T:309-310
            den0 = 1.0 + _scalar(mom.T.dot(u0))
            pw = w_new.copy()
            for _f in range(n_fixed):                                            # T:313-327
                _, p, v, g = fisher_metric(xx, pw, alpha)
                u = np.linalg.solve(g, mom)
                pw = w_new + (sgn * (step_size / 2) * nu) * u0 / den0 + (sgn * (step_size / 2) * nu) * u / (1.0 + _scalar(mom.T.dot(u)))
            w_new = pw
            _, p, v, g = fisher_metric(xx, w_new, alpha)                         # T:331-340
            inv_g = np.linalg.inv(g)
            inv_g_dg, tr = metric_partials(xx, p, v, inv_g)                      # T:345-356
            grad = likelihood_gradient(xx, t, w_new, alpha)                      # T:363-364
            u = inv_g.dot(mom)                                                   # T:368-373
            den = 1.0 + _scalar(mom.T.dot(u))
            last = np.array([[(nu / 2) * _scalar(mom.T.dot(inv_g_dg[k]).dot(u)) / den] for k in range(d)])
            mom = mom + sgn * (step_size / 2) * (grad - 0.5 * tr + last)
            rec["theta_steps"].append(w_new[:, 0].copy())
        prop_ljl = log_joint(xx, t, w_new, alpha)                                # T:379-383
        prop_logdet = np.sum(np.log(np.diag(np.linalg.cholesky(g))))             # T:385
        prop_h = -prop_ljl + prop_logdet + (nu / 2) * np.log(1.0 + _scalar(mom.T.dot(inv_g).dot(mom)))   # T:386
        cur_logdet = np.sum(np.log(np.diag(orig_chol)))                          # T:390
        cur_h = -cur_ljl + cur_logdet + (nu / 2) * np.log(1.0 + _scalar(mom0.T.dot(orig_inv_g).dot(mom0)))  # T:392
        ratio = _scalar(-prop_h + cur_h)                                         # T:395
        used_uniform = False
        take = bool(ratio > 0)
        if not take:
            used_uniform = True
            take = bool(ratio > np.log(tape.u_acc[it]))                          # T:398
        if take:
            cur_ljl, w = prop_ljl, w_new
            accepted[it] = True
        if record:
            rec.update({"mom_end": mom[:, 0].copy(), "theta_end": w_new[:, 0].copy(), "h_current": _scalar(cur_h),
                        "h_proposed": _scalar(prop_h), "ratio": ratio, "accepted": take, "used_uniform": used_uniform})
            records.append(rec)
        if it >= burn_in:
            samples[it - burn_in, :] = w.T                                       # T:403-405
    return samples, {"w": w[:, 0].copy(), "accepted": accepted, "steps": steps_taken, "records": records}


# --------------------------------------------------------------------------- IWLS (code/iwls.py)
def iwls_chain(xx, t, tape: DrawTape, n_iter=10000, burn_in=5000, alpha=ALPHA, record=False, w0=None):
    """One iterated-weighted-least-squares Metropolis chain under a draw tape; restates ``code/iwls.py:13-89``.

    The reference draws its proposal with ``np.random.multivariate_normal(current_mean, current_cov)`` (iwls.py:45),
    whose SVD factor depends on LAPACK sign conventions; under a tape that call is replaced by
    ``current_mean + cholesky(current_cov) @ z[it]`` (the same distribution) -- in the live reference run by patching
    that one NumPy attribute (oracle/ref_live.py: run_iwls), here directly.  ``u_acc[it]`` feeds
    ``np.random.uniform()`` (iwls.py:76), consumed only when ``ratio > 0`` is false.  Samples: row ``i - burn_in`` for
    ``i >= burn_in`` (iwls.py:84-85).  Everything else keeps the reference's NumPy operation order (bit-identical).
    """
    n_samples, dim = xx.shape
    beta = np.zeros(dim) if w0 is None else np.array(w0, dtype=float).reshape(dim)                # iwls.py:18
    beta_saved = np.zeros((n_iter - burn_in, dim))

    def joint(b):                                                                                 # iwls.py:22-25, :48-51
        log_prior = log_norm_pdf(np.zeros((1, dim)), b[:, np.newaxis], alpha)
        f = np.dot(xx, b)
        return np.dot(f.T, t) - np.sum(np.log(1 + np.exp(f))) + log_prior

    def moments(b):                                                                               # iwls.py:28-35, :54-61
        p = 1 / (1 + np.exp(-np.sum(b * xx, axis=1)))
        w = p * (np.ones(n_samples) - p)
        cov = np.linalg.inv(np.eye(dim) / alpha + (xx.T * np.tile(w[:, np.newaxis].T, (dim, 1))).dot(xx))
        # inv_W.dot(t - p) with inv_W = eye / W: a diagonal matrix product, i.e. (1 / W_i) (t - p)_i exactly
        z = xx.dot(b)[:, np.newaxis] + (1.0 / w)[:, np.newaxis] * (t - p[:, np.newaxis])
        mean = cov.dot(xx.T.dot(w[:, np.newaxis] * z))[:, 0]                                      # np.diag(W).dot(z)
        return cov, mean

    current_ljl = joint(beta)
    current_cov, current_mean = moments(beta)
    accepted = np.zeros(n_iter, dtype=bool)
    records = []
    for i in range(n_iter):
        beta_new = current_mean + np.linalg.cholesky(current_cov).dot(tape.z[i])                  # iwls.py:45 (see above)
        proposed_ljl = joint(beta_new)
        new_cov, new_mean = moments(beta_new)
        p_new_old = -np.sum(np.diag(np.log(np.linalg.cholesky(current_cov + np.eye(dim) * 1e-6))))   # iwls.py:64-66
        p_new_old -= 0.5 * (beta_new - current_mean).T.dot(np.linalg.inv(current_cov)).dot(beta_new - current_mean)
        p_old_new = -np.sum(np.diag(np.log(np.linalg.cholesky(new_cov + np.eye(dim) * 1e-6))))       # iwls.py:68-70
        p_old_new -= 0.5 * (beta - new_mean).T.dot(np.linalg.inv(new_cov)).dot(beta - new_mean)
        ratio = proposed_ljl + p_old_new - current_ljl - p_new_old                                # iwls.py:76
        used_uniform = False
        take = bool(ratio > 0)
        if not take:
            used_uniform = True
            take = bool(ratio > np.log(tape.u_acc[i]))
        if record:
            records.append({"theta": beta_new.copy(), "ratio": _scalar(ratio), "accepted": take, "used_uniform": used_uniform})
        if take:                                                                                  # iwls.py:76-81
            accepted[i] = True
            beta, current_cov, current_mean, current_ljl = beta_new, new_cov, new_mean, proposed_ljl
        if i >= burn_in:
            beta_saved[i - burn_in] = beta                                                        # iwls.py:84-85
    return beta_saved, {"w": beta.copy(), "log_joint": _scalar(current_ljl), "accepted": accepted, "records": records}


def iwls_chains(xx, t, tapes: list[DrawTape], **kw):
    out = [iwls_chain(xx, t, tp, **kw) for tp in tapes]
    return np.stack([o[0] for o in out]), [o[1] for o in out]


def mmala_chains(xx, t, tapes: list[DrawTape], **kw):
    out = [mmala_chain(xx, t, tp, **kw) for tp in tapes]
    return np.stack([o[0] for o in out]), [o[1] for o in out]


# --------------------------------------------------------------------------- batched helpers
def rmhmc_chains(xx, t, tapes: list[DrawTape], **kw):
    """Run independent chains one after another; returns samples (C, S, D) and the infos."""
    out = [rmhmc_chain(xx, t, tp, **kw) for tp in tapes]
    return np.stack([o[0] for o in out]), [o[1] for o in out]


def hmc_chains(xx, t, tapes: list[DrawTape], **kw):
    out = [hmc_chain(xx, t, tp, **kw) for tp in tapes]
    return np.stack([o[0] for o in out]), [o[1] for o in out]


def min_ess_total(chains: np.ndarray) -> float:
    """min_d sum_c ESS_{c,d} for chains (C, S, D) -- the bench metric's numerator."""
    total = np.zeros(chains.shape[2])
    for c in range(chains.shape[0]):
        total += ess(chains[c], chains.shape[1] - 1)[:, 0]
    return float(total.min())
