#!/bin/bash
# usage: probe_variant.sh <lib.so> <shape> <chains>   -- runs perf_probe with an alternative build of the library
cp riemannhamiltonianmontecarlo_b200/librmhmc_b200.so /tmp/lib_orig.so
cp $1 riemannhamiltonianmontecarlo_b200/librmhmc_b200.so
touch riemannhamiltonianmontecarlo_b200/librmhmc_b200.so
python scripts/perf_probe.py $2 $3 2>&1 | tail -7
cp /tmp/lib_orig.so riemannhamiltonianmontecarlo_b200/librmhmc_b200.so
