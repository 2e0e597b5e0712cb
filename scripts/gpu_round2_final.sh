#!/bin/bash
# round 2 final record: GPU test suite, bench lines of every sampler / workload, ncu launch list of the default command
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_final.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1), "roof", d["roofline"] and (d["roofline"]["kernel"][:24], round(d["roofline"]["frac"],3)))
    print("   ", {k:(round(v["ms_avg"],4), v.get("frac_of_peak") and round(v["frac_of_peak"],3)) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
run bench_r02_final_default
run bench_r02_final_hmc --steps 6 --warmup 3 --sampler hmc --no-cpu-baseline
run bench_r02_final_mmala --steps 6 --warmup 3 --sampler mmala --no-cpu-baseline
run bench_r02_final_australian --steps 10 --warmup 3 --workload australian --no-cpu-baseline
run bench_r02_final_dmma --steps 6 --warmup 3 --metric dmma --no-e2e --no-cpu-baseline
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/bench_r02_final_reference.err; tail -c 600 gpurun_out/bench_r02_final_reference.json
CMD="python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ncu_r02_final_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
