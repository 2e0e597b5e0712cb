#!/bin/bash
# round 2: launch list of the bench command + one `ncu --set full` capture per kernel of the INT8 / FP64 round
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
for triple in "i8_gemm_metric k_i8_gemm 3" "i8_gemm_leverage k_i8_gemm 1" "i8_vslice_iterate k_i8_vslice_mma< 2" "i8_vslice_closing k_i8_vslice_mma_closing 1" "mom_fixed_point k_pass<\(int\)3 1" "chain_solve k_chain_solve 3" "chain_factor k_chain_factor 1" "i8_qdigits k_i8_qdigits 1"; do
  set -- $triple
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
done
ls -la gpurun_out/*_raw.csv
