#!/bin/bash
# narrow last column chunk in k_i8_gemm: exactness self-test (incl. timing at 65 536 chains), then the bench A/B
mkdir -p gpurun_out
timeout 300 tests/native/i8_selftest 0 > gpurun_out/i8_selftest_narrow.log 2>&1; echo "selftest rc=$?"; grep -i "ms\|FAILED\|OK\|err" gpurun_out/i8_selftest_narrow.log | tail -14
TAG=episkip bash scripts/gpu_round2_ab.sh
