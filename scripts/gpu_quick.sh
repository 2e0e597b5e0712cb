#!/bin/bash
PROBE_PARTIALS=matrix_free timeout 600 python scripts/perf_probe.py german 65536 2>&1 | grep -v "iters"
