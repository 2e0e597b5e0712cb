#!/bin/bash
# round 2: small-batch regime of the momentum fixed point: k_mom_fp vs k_pass<MOMFP, 2 warps> (RMHMC_MOMFP_SMALL_PASS=1)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "rhat", d["rhat_max"], "acc", round(d["accept_rate"],3))
    print("   ", {k:(round(v["ms_avg"],4), round(v["share_of_step"],3), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for c in 8192 16384; do
run bench_small0_c$c --steps 10 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
RMHMC_MOMFP_SMALL_PASS=1 run bench_small1_c$c --steps 10 --warmup 3 --chains $c --no-e2e --no-cpu-baseline
done
run bench_small0_aus --steps 10 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
RMHMC_MOMFP_SMALL_PASS=1 run bench_small1_aus --steps 10 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
