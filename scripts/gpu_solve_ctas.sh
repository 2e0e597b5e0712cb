#!/bin/bash
# k_chain_solve<25> at different register caps (RMHMC_SOLVE_CTAS resident warps per SM), stand-alone (tests/native/tpc_bench.cu)
mkdir -p gpurun_out
for n in 16 20 24 32; do echo "== RMHMC_SOLVE_CTAS=$n"; timeout 120 tests/native/tpc_bench_t8_s$n 65536 25 | grep "^solve"; timeout 60 tests/native/tpc_bench_t8_s$n 8192 25 | grep "^solve"; done 2>&1 | tee gpurun_out/solve_ctas.log
