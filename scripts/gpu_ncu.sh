#!/bin/bash
# ncu passes (one gpurun call): launch list, then full captures of the hot kernels.
# The same command runs first WITHOUT ncu and must exit 0.  Reports are exported to CSV on the box
# (gpurun_out/ is capped at 64 MiB) and only the partials-build report itself is kept.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
capture() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
}
capture tbuild k_tbuild 3
capture metric_fp "k_metric.*Li0E" 8
capture metric_closing "k_metric.*Li1E" 3
capture chain_turn k_chain_turn 4
capture chain_solve k_chain_solve 8
capture chain_factor k_chain_factor 3
cp /tmp/prof_tbuild.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
