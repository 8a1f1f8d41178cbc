#!/bin/bash
# ncu --set full of the two tau setup kernels (one launch each) inside a CNRK2 step at C4
mkdir -p gpurun_out
T=r02ab
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tau_factor|tau_profiles" -c 2 -o gpurun_out/${T}_tausetup python bench.py --workload c4 --stepper cnrk2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/${T}_tausetup.ncu-rep --page raw --csv > gpurun_out/${T}_tausetup_raw.csv 2>/dev/null
ls -la gpurun_out/${T}_*
