#!/bin/bash
# ncu captures (full set, with source) of the restructured k_pass<MOMFP, 8> / k_pass<PAIR, 8>; the command has exited 0 without ncu first.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -20 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip
  local name=$1 rx=$2 skip=$3
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -o /tmp/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/ncu_r02_${name}_details.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/ncu_r02_${name}_source.csv 2>/dev/null
}
cap final2_mom_fixed_point '^k_pass$' 1
cap final2_pair_pass '^k_pass$' 2
ls -la gpurun_out | tail -8
