#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -20 gpurun_out/plain.log; exit 1; }
for triple in "i8_vslice_iterate ^k_i8_vslice_mma$ 2" "mom_fixed_point ^k_pass$ 1"; do
  set -- $triple
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
done
ls -la gpurun_out/i8_vslice_iterate_raw.csv gpurun_out/mom_fixed_point_raw.csv
