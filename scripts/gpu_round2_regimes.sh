#!/bin/bash
# round 2: per-kernel times of the SMALL (auto below 18 944 chains) and LARGE launch regimes at strong-scaling batch sizes
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2))
    print("   ", {k:round(v["ms_avg"],4) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for c in 4096 8192 12288 16384; do
RMHMC_MOMFP_SMALL_PASS=1 run regime_auto_c$c --steps 6 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
RMHMC_LAUNCH_REGIME=large run regime_large_c$c --steps 6 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
done
