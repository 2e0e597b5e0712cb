#!/bin/bash
# usage: gpu_ncu_multi.sh "<tag> <kernel regex> <skip>" ...   -- one full-set capture per triple, CSV export on the box
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline ${BENCH_EXTRA}"
$CMD > gpurun_out/plain.log 2>&1 || { tail -20 gpurun_out/plain.log; exit 1; }
for triple in "$@"; do
  set -- $triple
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
