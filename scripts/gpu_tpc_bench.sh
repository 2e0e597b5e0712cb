#!/bin/bash
# stand-alone timing / agreement check of the thread-per-chain Cholesky kernels (tests/native/tpc_bench.cu)
mkdir -p gpurun_out
for b in tests/native/tpc_bench_t*; do echo "== $b"; timeout 120 $b 65536 25; done 2>&1 | tee gpurun_out/tpc_bench.log
timeout 120 tests/native/tpc_bench_t8 8192 25 2>&1 | tee -a gpurun_out/tpc_bench.log
timeout 120 tests/native/tpc_bench_t8 4099 15 2>&1 | tee -a gpurun_out/tpc_bench.log
timeout 120 tests/native/tpc_bench_t8 1000 32 2>&1 | tee -a gpurun_out/tpc_bench.log
timeout 120 tests/native/tpc_bench_t8 1000 7 2>&1 | tee -a gpurun_out/tpc_bench.log
