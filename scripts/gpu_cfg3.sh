#!/bin/bash
# BASELINE.json configs[2]: synthetic N=100k, D=100, 1024 chains (one B200): per-kernel timing of a few rounds
mkdir -p gpurun_out
PROBE_ROUNDS=3 PROBE_WARM=1 PROBE_PARTIALS=matrix_free timeout 1500 python scripts/perf_probe.py synthetic:100000:100:1236 1024 2>&1 | tee gpurun_out/probe_cfg3.log
