#!/bin/bash
# y-transform as shared-memory FFT (yfft.cu) against the DMMA contraction: stage times at C4, parity subset
mkdir -p gpurun_out
T=r02ae
for v in "0 8" "1 8" "1 4"; do
  set -- $v
  CF_YFFT=$1 CF_YFFT_C=$2 timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_c4_yfft$1_c$2.json 2> gpurun_out/${T}_bench_c4_yfft$1_c$2.err
  python - <<P
import json
l=json.loads(open('gpurun_out/${T}_bench_c4_yfft$1_c$2.json').read().strip().splitlines()[-1])
print("yfft=$1 C=$2", l['ms_per_step'], {k: round(v,3) for k,v in l.get('stages_ms',{}).items()} if 'stages_ms' in l else [k for k in l.keys()])
P
done
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "transform or nonlinear or step or golden or stepper or c1 or layout or graph" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
