#!/bin/bash
mkdir -p gpurun_out
for nt in 96 128 160 192; do
  CF_ZP_THREADS=$nt timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02v_zp$nt.json 2> gpurun_out/r02v_zp$nt.err
done
python scripts/print_bench.py gpurun_out/r02v_zp*.json | cut -c1-260
