#!/bin/bash
# round 2: k_pass / k_hmc_rounds with compile-time tile counts: parity tests, RMHMC + HMC bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pass.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_pass.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "rhat", d["rhat_max"], "acc", round(d["accept_rate"],3), "e2e", d["e2e"] and round(d["e2e"]["value"]), "roof", d["roofline"] and (d["roofline"]["kernel"][:24], round(d["roofline"]["frac"],3)), "peaks", d["peaks"])
    print("   ", {k:(round(v["ms_avg"],4), round(v["share_of_step"],3), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
run bench_pass_c65536 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline
run bench_pass_c8192 --steps 10 --warmup 3 --chains 8192 --no-e2e --no-cpu-baseline
run bench_pass_hmc --steps 6 --warmup 3 --sampler hmc --no-e2e --no-cpu-baseline
run bench_pass_aus --steps 10 --warmup 3 --workload australian --no-e2e --no-cpu-baseline
