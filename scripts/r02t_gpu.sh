#!/bin/bash
mkdir -p gpurun_out
for kb in 112 56 30; do
  CF_TAU_TILE_KB=$kb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02t_bench_tile$kb.json 2> gpurun_out/r02t_bench_tile$kb.err; echo "tile $kb exit $?"
done
python scripts/print_bench.py gpurun_out/r02t_bench_tile*.json
CF_TAU_TILE_KB=56 timeout 600 python -m pytest tests/test_gpu.py -q -k "tausolve or steppers or golden or c1_one" 2>&1 | tail -2
