#!/bin/bash
# k_i8_gemm changes: exactness self-test (tests/native/i8_selftest.cu, built beforehand), then the bench A/B at 65 536 and 8192 chains
mkdir -p gpurun_out
timeout 300 tests/native/i8_selftest 0 > gpurun_out/i8_selftest_narrow.log 2>&1; echo "selftest rc=$?"; grep -i "ms\|FAILED\|OK\|err" gpurun_out/i8_selftest_narrow.log | tail -14
TAG=episkip bash scripts/gpu_round2_ab.sh
