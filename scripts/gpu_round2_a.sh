#!/bin/bash
# round 2: tests + bench lines of every sampler + per-GPU share of the strong-scaled configs[3]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
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
run bench_r02_default --steps 10 --warmup 3
run bench_r02_c8192 --steps 10 --warmup 3 --chains 8192 --no-e2e --no-cpu-baseline
run bench_r02_c16384 --steps 10 --warmup 3 --chains 16384 --no-e2e --no-cpu-baseline
run bench_r02_c32768 --steps 10 --warmup 3 --chains 32768 --no-e2e --no-cpu-baseline
run bench_r02_hmc --steps 6 --warmup 3 --sampler hmc --no-cpu-baseline
run bench_r02_mmala --steps 6 --warmup 3 --sampler mmala --no-cpu-baseline
run bench_r02_australian --steps 10 --warmup 3 --workload australian --no-cpu-baseline
