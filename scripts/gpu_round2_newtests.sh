#!/bin/bash
# the tests added last (dimension sweep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_dimension_sweep" --durations=3 > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_new.log
