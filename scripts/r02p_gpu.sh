#!/bin/bash
# closing 1-GPU run: full parity suite, default bench line (+ reference arm), launch list + ncu of the z-pass
mkdir -p gpurun_out
T=${1:-r02p}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
grep -E "variable dt:|findsoln:|passed|failed|exit" gpurun_out/${T}_pytest_gpu.log | cut -c1-600 | tail -6
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_default.json 2>&1 | tail -3
cut -c1-300 gpurun_out/${T}_bench_reference.json
KERNELS="zpass" bash scripts/gpu_profile.sh ${T} > gpurun_out/${T}_profile.log 2>&1
python scripts/ncu_summary.py "gpurun_out/${T}_raw_*.csv" > gpurun_out/${T}_ncu_summary.md 2>&1 || true
grep -E "time_duration|inst_executed|issue_active" gpurun_out/${T}_ncu_summary.md | head -4
