#!/bin/bash
# HMC: parity of the fused many-rounds-per-launch kernel, then the bench with both launch shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "hmc" > gpurun_out/hmc_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/hmc_tests.log
tail -5 gpurun_out/hmc_tests.log
timeout 600 python bench.py --sampler hmc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hmc_fused.json 2> gpurun_out/bench_hmc_fused.err; echo "fused rc=$?"
RMHMC_HMC_FUSED=0 timeout 600 python bench.py --sampler hmc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hmc_unfused.json 2> gpurun_out/bench_hmc_unfused.err; echo "unfused rc=$?"
tail -c 1500 gpurun_out/bench_hmc_fused.json; echo; tail -c 600 gpurun_out/bench_hmc_unfused.json
