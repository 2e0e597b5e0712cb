"""Drive the UNMODIFIED reference under a host-supplied RNG tape.  TEST INFRASTRUCTURE ONLY.

Only usable where ``/root/reference`` is mounted (the build container); the GPU box
does not have it, so nothing that runs there imports this module.  It is used by
``tests/golden/make_golden.py`` to produce the committed fixtures and by the
``reference``-marked CPU tests that re-check the oracle against the live code.

The reference looks up ``np.random.randn`` / ``np.random.rand`` at call time
(rmhmc.py:80,89,90,181; hmc.py:41,48,78), so replacing those two attributes for the
duration of a call feeds it a :class:`oracle.blr_oracle.DrawTape` without touching
its source.  Per-step state is captured with ``sys.settrace`` on the sampler's code
object (line numbers are those of rmhmc.py / hmc.py at the pinned reference commit).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

import numpy as np

REFERENCE_CODE = "/root/reference/code"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, "rmhmc.py"))


def load_reference():
    """Import the reference's ``rmhmc``, ``hmc`` and ``tools`` modules as they are."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_CODE}")
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    mods = {name: importlib.import_module(name) for name in ("tools", "rmhmc", "hmc")}
    for name, mod in mods.items():
        assert os.path.dirname(os.path.abspath(mod.__file__)) == REFERENCE_CODE, (name, mod.__file__)
    return mods


class _TapeFeeder:
    """Stands in for np.random.randn / rand and replays one chain's DrawTape."""

    def __init__(self, tape, with_direction: bool):
        self.tape = tape
        self.with_direction = with_direction
        self.it = -1
        self.rand_calls_this_iter = 0
        self.uniform_used = []

    def randn(self, *shape):
        if len(shape) == 2:               # randn(1, D): a new iteration starts
            self.it += 1
            self.rand_calls_this_iter = 0
            self.uniform_used.append(False)
            return self.tape.z[self.it].reshape(shape).copy()
        assert shape == () and self.with_direction, "unexpected randn() call"
        return float(self.tape.z_dir[self.it])

    def rand(self, *shape):
        assert shape == ()
        self.rand_calls_this_iter += 1
        if self.rand_calls_this_iter == 1:
            return float(self.tape.u_step[self.it])
        assert self.rand_calls_this_iter == 2, "more than two rand() calls in one iteration"
        self.uniform_used[self.it] = True
        return float(self.tape.u_acc[self.it])


@contextlib.contextmanager
def _patched_rng(feeder):
    saved = (np.random.randn, np.random.rand)
    np.random.randn, np.random.rand = feeder.randn, feeder.rand
    try:
        yield
    finally:
        np.random.randn, np.random.rand = saved


def _run_traced(func, code_lines, args, kwargs):
    """Call ``func`` and snapshot selected locals whenever one of ``code_lines`` is about to run."""
    events = []
    code = func.__code__

    def local_tracer(frame, event, arg):
        if event == "line" and frame.f_lineno in code_lines:
            names = code_lines[frame.f_lineno]
            snap = {"line": frame.f_lineno}
            for nm in names:
                val = frame.f_locals.get(nm)
                snap[nm] = np.array(val, dtype=float).copy() if val is not None else None
            events.append(snap)
        return local_tracer

    def global_tracer(frame, event, arg):
        return local_tracer if frame.f_code is code else None

    sink = io.StringIO()
    sys.settrace(global_tracer)
    try:
        with contextlib.redirect_stdout(sink):
            out = func(*args, **kwargs)
    finally:
        sys.settrace(None)
    return out, events


def run_rmhmc(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size, n_fixed, trace=True):
    """Reference ``rmhmc.RMHMC`` under ``tape``.

    Returns ``(wSaved, info)`` where ``info['steps']`` lists, per leapfrog step, the
    position entering the closing half-step (line 134) and ``info['iters']`` the
    momentum after the trajectory, CurrentH, ProposedH, Ratio (line 181).
    """
    ref = load_reference()["rmhmc"]
    feeder = _TapeFeeder(tape, with_direction=True)
    lines = {
        134: ("IterationNum", "StepNum", "wNew"),                       # position after step (+hack)
        166: ("IterationNum", "ProposedMomentum", "wNew", "RandomStep", "TimeStep", "OriginalMomentum"),
        181: ("IterationNum", "CurrentH", "ProposedH", "Ratio"),
    } if trace else {}
    with _patched_rng(feeder):
        if trace:
            (w_saved, _), events = _run_traced(
                ref.RMHMC, lines, (xx, t),
                dict(NumOfIterations=n_iter, BurnIn=burn_in, NumOfLeapFrogSteps=n_leapfrog,
                     StepSize=step_size, NumOfNewtonSteps=n_fixed))
        else:
            with contextlib.redirect_stdout(io.StringIO()):
                w_saved, _ = ref.RMHMC(xx, t, NumOfIterations=n_iter, BurnIn=burn_in,
                                       NumOfLeapFrogSteps=n_leapfrog, StepSize=step_size,
                                       NumOfNewtonSteps=n_fixed)
            events = []
    info = {"uniform_used": np.array(feeder.uniform_used, dtype=bool),
            "steps": [e for e in events if e["line"] == 134],
            "ends": [e for e in events if e["line"] == 166],
            "iters": [e for e in events if e["line"] == 181]}
    return w_saved, info


def run_hmc(xx, t, tape, n_iter, burn_in, n_leapfrog, step_size):
    """Reference ``hmc.HMC`` under ``tape`` (no direction draw)."""
    ref = load_reference()["hmc"]
    feeder = _TapeFeeder(tape, with_direction=False)
    lines = {77: ("IterationNum", "CurrentH", "ProposedH", "Ratio", "wNew", "ProposedMomentum")}
    with _patched_rng(feeder):
        (w_saved, _), events = _run_traced(
            ref.HMC, lines, (xx, t),
            dict(NumOfIterations=n_iter, BurnIn=burn_in, NumOfLeapFrogSteps=n_leapfrog,
                 StepSize=step_size))
    return w_saved, {"uniform_used": np.array(feeder.uniform_used, dtype=bool), "iters": events}


def run_iwls(xx, t, tape, max_iter, burn_in, alpha=100):
    """Reference ``iwls.iwls`` (code/iwls.py) under ``tape``.

    ``np.random.multivariate_normal`` (iwls.py:45) is replaced for the call by ``mean + cholesky(cov) @ z[it]`` and
    ``np.random.uniform`` (iwls.py:76) by ``u_acc[it]``: every other line of the unmodified reference runs as is.
    Returns ``(beta_saved, info)`` with the per-iteration proposal and ratio (line 76).
    """
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    load_reference()
    ref = importlib.import_module("iwls")
    assert os.path.dirname(os.path.abspath(ref.__file__)) == REFERENCE_CODE
    state = {"it": -1, "used": []}

    def mvn(mean, cov):
        state["it"] += 1
        state["used"].append(False)
        return mean + np.linalg.cholesky(cov).dot(tape.z[state["it"]])

    def uniform():
        state["used"][state["it"]] = True
        return float(tape.u_acc[state["it"]])

    saved = (np.random.multivariate_normal, np.random.uniform)
    np.random.multivariate_normal, np.random.uniform = mvn, uniform
    try:
        (beta_saved, _), events = _run_traced(ref.iwls, {76: ("i", "beta_new", "ratio")}, (xx, t),
                                              dict(alpha=alpha, max_iter=max_iter, burn_in=burn_in))
    finally:
        np.random.multivariate_normal, np.random.uniform = saved
    return beta_saved, {"uniform_used": np.array(state["used"], dtype=bool), "iters": events}
