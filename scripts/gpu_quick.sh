#!/bin/bash
# one C4 bench line (device-resident only) + the GPU parity tests
mkdir -p gpurun_out
scripts/gpu_ab.sh X ${1:-run}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
