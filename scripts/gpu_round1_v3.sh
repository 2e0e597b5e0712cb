#!/bin/bash
# Round-1 record job (v3, matrix-free partials): parity tests, both bench arms, both partials modes, cfg2, ncu launch
# list and full-set captures of every kernel class (CSV exported on the box; gpurun_out is capped at 64 MiB).
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; tail -c 600 gpurun_out/bench_b200.json
python bench.py --partials tensor --no-cpu-baseline > gpurun_out/bench_b200_tensor.json 2> gpurun_out/bench_tensor.err; tail -c 300 gpurun_out/bench_b200_tensor.json
python bench.py --workload australian --no-cpu-baseline > gpurun_out/bench_b200_australian4096.json 2> gpurun_out/bench_aus.err; tail -c 300 gpurun_out/bench_b200_australian4096.json
CMD="python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
capture() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
}
capture metric_fp k_metric 8
capture metric_closing k_metric 12
capture mom_fixed_point k_pass 3
capture pair_pass k_pass 4
capture leverage_gemm k_tbuild_pre 2
capture chain_solve k_chain_solve 8
capture chain_factor k_chain_factor 3
capture mf_turn k_mf_turn 3
du -sh gpurun_out
