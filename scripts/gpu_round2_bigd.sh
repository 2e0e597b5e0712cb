#!/bin/bash
# 32 < D and > 16384 rows through the INT8 digit GEMM: parity tests, then BASELINE configs[2] with both metric modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "large_dim or row_sharded_code_path or int8_build_splits or metric_seam" > gpurun_out/bigd_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/bigd_tests.log
tail -15 gpurun_out/bigd_tests.log
timeout 600 python bench.py --workload cfg3 --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_cfg3_i8.json 2> gpurun_out/bench_cfg3_i8.err; echo "cfg3 i8 rc=$?"
timeout 600 python bench.py --workload cfg3 --metric dmma --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_cfg3_dmma.json 2> gpurun_out/bench_cfg3_dmma.err; echo "cfg3 dmma rc=$?"
tail -3 gpurun_out/bench_cfg3_i8.err
