#!/bin/bash
# round 2 (v8): full GPU test suite, then bench lines at the strong-scaling batch sizes and the default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_v8.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"]), "roof", d["roofline"] and (d["roofline"]["kernel"][:24], round(d["roofline"]["frac"],3)))
    print("   ", {k:round(v["ms_avg"],4) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for c in 8192 16384 32768; do run bench_v8_c$c --steps 6 --warmup 3 --chains $c --no-e2e --no-cpu-baseline; done
run bench_v8_default --steps 10 --warmup 3
run bench_v8_aus --steps 10 --warmup 3 --workload australian --no-cpu-baseline
