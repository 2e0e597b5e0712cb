#!/bin/bash
# round 2: k_i8_vslice_mma with 2 m-tiles per warp (default) vs 4 (RMHMC_VSLICE_MT=4)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "rhat", d["rhat_max"])
    print("   ", {k:(round(v["ms_avg"],4), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for c in 65536 8192; do
run vsmt2_c$c --steps 6 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
RMHMC_VSLICE_MT=4 run vsmt4_c$c --steps 6 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
done
run vsmt2_aus --steps 6 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
RMHMC_VSLICE_MT=4 run vsmt4_aus --steps 6 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
