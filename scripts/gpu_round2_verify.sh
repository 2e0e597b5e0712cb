#!/bin/bash
# final verification of the committed build: smoke + the whole GPU test suite + the default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/bench_r02_final_default.json 2> gpurun_out/bench_r02_final_default.err; tail -c 400 gpurun_out/bench_r02_final_default.json
