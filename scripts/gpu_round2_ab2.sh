#!/bin/bash
# quick A/B: default, HMC and Australian-shaped bench lines (no tests)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "rhat", d["rhat_max"], "roof", d["roofline"] and (d["roofline"]["kernel"][:24], round(d["roofline"]["frac"],3)))
    print("   ", {k:(round(v["ms_avg"],4), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
TAG=${TAG:-ab2}
run ${TAG}_default --steps 6 --warmup 3 --no-e2e --no-cpu-baseline
run ${TAG}_hmc --steps 6 --warmup 3 --sampler hmc --no-e2e --no-cpu-baseline
run ${TAG}_aus --steps 6 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
