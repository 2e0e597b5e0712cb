#!/bin/bash
# the environment-switched kernel variants under the parity tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_kernel_variants_behind_environment_switches" --durations=8 > gpurun_out/pytest_variants.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_variants.log
