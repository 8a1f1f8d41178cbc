#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02o}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 900 python scripts/bench_findsoln.py > gpurun_out/${T}_findsoln_c3.json 2> gpurun_out/${T}_findsoln_c3.err; echo "findsoln c3 exit $?"; cut -c1-400 gpurun_out/${T}_findsoln_c3.json; tail -3 gpurun_out/${T}_findsoln_c3.err
CFGPU_GRAPH=0 timeout 900 python scripts/bench_findsoln.py > gpurun_out/${T}_findsoln_c3_nograph.json 2> gpurun_out/${T}_findsoln_c3_nograph.err; echo "findsoln c3 (no graph) exit $?"; cut -c1-300 gpurun_out/${T}_findsoln_c3_nograph.json
CFGPU_NO_POOL=1 timeout 900 python scripts/bench_findsoln.py > gpurun_out/${T}_findsoln_c3_nopool.json 2> gpurun_out/${T}_findsoln_c3_nopool.err; echo "findsoln c3 (no pool) exit $?"; cut -c1-300 gpurun_out/${T}_findsoln_c3_nopool.json
timeout 600 python bench.py --steps 5 --warmup 3 --stepper cnrk2 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_cnrk2.json 2> gpurun_out/${T}_bench_cnrk2.err; echo "bench cnrk2 exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_cnrk2.json
