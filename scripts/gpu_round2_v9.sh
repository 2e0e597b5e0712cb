#!/bin/bash
# round 2 (v9): conflict-free row pairing in k_pass / k_hmc_rounds: GPU tests, bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_v9.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"]), "roof", d["roofline"] and (d["roofline"]["kernel"][:24], round(d["roofline"]["frac"],3)))
    print("   ", {k:(round(v["ms_avg"],4), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
run bench_v9_c65536 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline
run bench_v9_c8192 --steps 6 --warmup 3 --chains 8192 --no-e2e --no-cpu-baseline
run bench_v9_hmc --steps 6 --warmup 3 --sampler hmc --no-e2e --no-cpu-baseline
