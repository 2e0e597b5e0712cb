#!/bin/bash
# Round-1 measurement job (run under gpurun): parity tests, both bench arms, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; tail -c 3000 gpurun_out/bench_b200.json; tail -5 gpurun_out/bench_b200.err
