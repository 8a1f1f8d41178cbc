#!/bin/bash
# closing 1-GPU run of round 2: smoke, full parity suite, default bench line + reference arm, ncu launch list at the same HEAD
mkdir -p gpurun_out
T=${1:-r02z}
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
grep -E "converged:|stopped:|passed|failed|exit" gpurun_out/${T}_pytest_gpu.log | cut -c1-300 | tail -5
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_default.json 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1; echo "launch list exit $?"
for wl in c1 c2 c5 golden; do timeout 600 python bench.py --workload $wl --steps 300 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_${wl}.json 2> gpurun_out/${T}_bench_${wl}.err; done
python scripts/print_bench.py gpurun_out/${T}_bench_c1.json gpurun_out/${T}_bench_c2.json gpurun_out/${T}_bench_c5.json gpurun_out/${T}_bench_golden.json 2>&1 | cut -c1-250
