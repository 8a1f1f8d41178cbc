#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02m}
timeout 900 python -m pytest tests/test_gpu.py -q -s -k "symmetry or findsoln or graph" > gpurun_out/${T}_pytest_subset.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_subset.log
grep -E "findsoln:|graph replay|passed|failed|exit|Error|error" gpurun_out/${T}_pytest_subset.log | cut -c1-1500 | tail -12
timeout 900 python scripts/bench_findsoln.py > gpurun_out/${T}_findsoln_c3.json 2> gpurun_out/${T}_findsoln_c3.err; echo "findsoln c3 exit $?"; cat gpurun_out/${T}_findsoln_c3.json | cut -c1-1200; tail -3 gpurun_out/${T}_findsoln_c3.err
CFGPU_GRAPH=0 timeout 900 python scripts/bench_findsoln.py > gpurun_out/${T}_findsoln_c3_nograph.json 2> gpurun_out/${T}_findsoln_c3_nograph.err; echo "findsoln c3 (no graph) exit $?"; cut -c1-300 gpurun_out/${T}_findsoln_c3_nograph.json
