#!/bin/bash
# half-length FFT y-transform: stage times at C4 for C = 8 / 16 / 4 and the full-length variant, parity subset
mkdir -p gpurun_out
T=${1:-r02aj}
for v in ${VARIANTS:-"1 8"}; do
  set -- $v
  CF_YFFT_HALF=$1 CF_YFFT_C=$2 timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_c4_half$1_c$2.json 2> gpurun_out/${T}_bench_c4_half$1_c$2.err
  python - <<P
import json
l=json.loads(open('gpurun_out/${T}_bench_c4_half$1_c$2.json').read().strip().splitlines()[-1])
print("half=$1 C=$2", l['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in l['stages'].items()})
P
done
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "transform or nonlinear or step or golden or stepper or c1 or layout or graph" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
