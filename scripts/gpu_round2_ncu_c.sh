#!/bin/bash
# ncu captures of the kernels added late in round 2: k_hmc_rounds (German-shaped, 65536 chains) and k_i8_gemm / k_i8_vdigits at
# the configs[2] shape (N = 1e5, D = 100, 1024 chains).  Each command has exited 0 without ncu first.
mkdir -p gpurun_out
CMD1="python bench.py --sampler hmc --steps 1 --warmup 3 --rounds-per-step 8 --no-e2e --no-cpu-baseline"
CMD2="python bench.py --workload cfg3 --steps 1 --warmup 3 --rounds-per-step 1 --no-e2e --no-cpu-baseline"
$CMD1 > gpurun_out/plain1.log 2>&1 || { tail -20 gpurun_out/plain1.log; exit 1; }
$CMD2 > gpurun_out/plain2.log 2>&1 || { tail -20 gpurun_out/plain2.log; exit 1; }
cap() {  # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/ncu_r02_${name}_details.csv 2>/dev/null
}
cap hmc_rounds '^k_hmc_rounds$' 1 $CMD1
cap cfg3_i8_gemm '^k_i8_gemm$' 3 $CMD2
cap cfg3_i8_vdigits '^k_i8_vdigits$' 3 $CMD2
ls -la gpurun_out/ncu_r02_hmc_rounds_raw.csv gpurun_out/ncu_r02_cfg3_i8_gemm_raw.csv gpurun_out/ncu_r02_cfg3_i8_vdigits_raw.csv
