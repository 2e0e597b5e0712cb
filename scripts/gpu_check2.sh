#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
PROBE_PARTIALS=matrix_free timeout 600 python scripts/perf_probe.py australian 4096 2>&1 | grep "partials=\|quad_pass\|trace_pass"
for fm in 1 2; do
echo "== RMHMC_FUSE_MOMENTUM=$fm"
RMHMC_FUSE_MOMENTUM=$fm timeout 900 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_fm$fm.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fm$fm.json').read().strip().splitlines()[-1])
print(d['value'], d['leapfrog_steps_per_sec'], d['ms_per_step'])
for k,v in d['kernels'].items(): print('   ',k, round(v['ms_avg'],4), v['launches'], round(v['share_of_step'],3))
PY
done
