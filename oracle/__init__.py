"""CPU oracle for the batched RMHMC/HMC hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU arm being timed.
The product package (``riemannhamiltonianmontecarlo_b200``) never imports it and
fails loudly when its CUDA library is missing.

Parity pin: the restatement in :mod:`oracle.blr_oracle` is validated bit-for-bit
against the unmodified reference (``/root/reference/code/{rmhmc,hmc,tools}.py``)
run under a host-supplied RNG tape (:mod:`oracle.ref_live`); the resulting vectors
are committed under ``tests/golden/`` by ``tests/golden/make_golden.py``.  The
reference ships no tests or golden vectors of its own.
Exception: ``mmala_chain`` ports MATLAB-only code (BLR_mMALA.m) that cannot be run here -- parity unpinned.
"""
