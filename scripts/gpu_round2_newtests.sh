#!/bin/bash
# the tests added last (German-shaped HMC fixture, mid-scale German-shaped batch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_hmc_matches_reference or test_hmc_large_regime or test_mid_scale" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_new.log
