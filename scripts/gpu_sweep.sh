#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  echo "=== dbg=$dbg"
  RMHMC_DBG=$dbg PROBE_PARTIALS=matrix_free PROBE_ROUNDS=4 PROBE_WARM=2 timeout 600 python scripts/perf_probe.py german 65536 2>&1 | grep "metric_fp\|metric_clos"
done 2>&1 | tee gpurun_out/sweep_dbg.log
