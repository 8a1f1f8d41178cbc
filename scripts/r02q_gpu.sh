#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02q}
timeout 1500 python -m pytest tests/test_refprogs.py -m gpu -q -s -k "simulateflow or benchmark or time_integration or couette or orr" > gpurun_out/${T}_pytest_refprogs.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_refprogs.log
grep -E "simulateflow C1|reference benchmark|passed|failed|exit|Error" gpurun_out/${T}_pytest_refprogs.log | cut -c1-700 | tail -8
