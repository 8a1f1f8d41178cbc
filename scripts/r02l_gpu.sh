#!/bin/bash
# z-pass f_z pairing + CUDA-graph replay: parity subset, C4 bench, small-grid benches with and without graph replay
mkdir -p gpurun_out
T=${1:-r02l}
timeout 900 python -m pytest tests/test_gpu.py -q -s -k "graph or nonlinear or steppers or golden or full_size or c1_one or tile_layout" > gpurun_out/${T}_pytest_subset.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_subset.log
grep -E "graph replay|passed|failed|exit" gpurun_out/${T}_pytest_subset.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench c4 exit $?"
for wl in c1 golden c2 c5; do
  for g in 0 1; do
    CFGPU_GRAPH=$g timeout 600 python bench.py --workload $wl --steps 300 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_${wl}_graph$g.json 2> gpurun_out/${T}_bench_${wl}_graph$g.err; echo "bench $wl graph=$g exit $?"
  done
done
python scripts/print_bench.py gpurun_out/${T}_bench_*.json 2>&1 | tail -12
for f in gpurun_out/${T}_bench_*.err; do tail -n 2 $f; done
