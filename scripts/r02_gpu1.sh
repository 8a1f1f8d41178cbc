#!/bin/bash
# 1-GPU round: parity tests, bench A/B of the y-GEMM variants (CF_YG_SPLIT), RK-type stepper stage times
mkdir -p gpurun_out
T=${1:-r02b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
for v in ${YGV:-0 1 2}; do
CF_YG_SPLIT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_split$v.json 2> gpurun_out/${T}_bench_split$v.err; echo "bench split$v exit $?"
done
timeout 600 python bench.py --steps 5 --warmup 3 --stepper cnrk2 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_cnrk2.json 2> gpurun_out/${T}_bench_cnrk2.err; echo "bench cnrk2 exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_split*.json gpurun_out/${T}_bench_cnrk2.json 2>&1 | tail -40
for f in gpurun_out/${T}_bench_*.err; do tail -n 3 $f; done
