#!/bin/bash
# parity tests + per-kernel timing probe of both partials modes (run under gpurun)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
PROBE_PARTIALS=${PROBE_PARTIALS:-matrix_free} timeout 600 python scripts/perf_probe.py german 65536 > gpurun_out/probe_german.log 2>&1; cat gpurun_out/probe_german.log
PROBE_PARTIALS=${PROBE_PARTIALS:-matrix_free} timeout 600 python scripts/perf_probe.py australian 4096 > gpurun_out/probe_aus.log 2>&1; cat gpurun_out/probe_aus.log
