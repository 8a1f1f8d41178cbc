#!/bin/bash
# ncu --set full of the FFT y-transform kernels (inverse and forward launch of one step at C4)
mkdir -p gpurun_out
T=${1:-r02af}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"yfft" -s 4 -c 2 -o gpurun_out/${T}_yfft python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/${T}_yfft.ncu-rep --page raw --csv > gpurun_out/${T}_yfft_raw.csv 2>/dev/null
ls -la gpurun_out/${T}_*
