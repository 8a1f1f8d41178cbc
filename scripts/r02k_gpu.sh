#!/bin/bash
# 1-GPU: full parity suite (incl. the reference's own programs), default bench line, launch list, full ncu capture of the hot kernels
mkdir -p gpurun_out
T=${1:-r02k}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -6 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_default.json 2>&1 | tail -12
tail -n 3 gpurun_out/${T}_bench_default.err
KERNELS="${KERNELS:-zpass tau_solve ygemm xpass_inverse xpass_forward}" bash scripts/gpu_profile.sh ${T} > gpurun_out/${T}_profile.log 2>&1
python scripts/ncu_summary.py "gpurun_out/${T}_raw_*.csv" > gpurun_out/${T}_ncu_summary.md 2>&1 || true
tail -30 gpurun_out/${T}_ncu_summary.md
